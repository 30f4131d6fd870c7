mkdir -p gpurun_out
LEMO_PERFRAME_TL=1 timeout 300 python tools/run_stage.py perframe > gpurun_out/perframe_tl.log 2>&1; tail -4 gpurun_out/perframe_tl.log
( time timeout 1200 python -m pytest tests/test_gpu_body.py tests/test_gpu_fit.py tests/test_gpu_prox_loss.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/pytest_a.log 2>&1
tail -c 1500 gpurun_out/pytest_a.log
timeout 900 python bench.py --skip-cpu-baseline --skip-extra --min-seconds 0.3 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err; tail -2 gpurun_out/bench_pf.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pf.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'],'perframe',d.get('perframe'),'infill',{k:v for k,v in (d.get('infill') or {}).items() if 'ms' in k},'prox',{k:v for k,v in (d.get('prox') or {}).items() if 'ms_per' in k})
PY
