mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_ae.py tests/test_gpu_infill.py tests/test_gpu_fit.py tests/test_scripts_surface.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/pytest_a.log 2>&1
tail -c 600 gpurun_out/pytest_a.log
timeout 900 python bench.py --skip-cpu-baseline --skip-extra --skip-prox --skip-perframe --min-seconds 0.3 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err; tail -2 gpurun_out/bench_pf.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pf.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'],'infill',{k:v for k,v in (d.get('infill') or {}).items() if 'ms' in k})
PY
