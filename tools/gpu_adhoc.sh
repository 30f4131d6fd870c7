mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_body.py tests/test_gpu_fit.py tests/test_gpu_prox_loss.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/pytest_a.log 2>&1
tail -c 400 gpurun_out/pytest_a.log
timeout 120 python tools/diag_lbs.py 120 > gpurun_out/diag_lbs.log 2>&1; head -2 gpurun_out/diag_lbs.log
timeout 900 python bench.py --skip-cpu-baseline --skip-extra --skip-infill --min-seconds 0.5 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err; tail -2 gpurun_out/bench_pf.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pf.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'],'lbs',d['roofline_lbs']['ms'],d['roofline_lbs']['frac'],'perframe',d['perframe']['us_per_iteration'],'prox',{k:v for k,v in (d.get('prox') or {}).items() if 'ms_per' in k})
PY
