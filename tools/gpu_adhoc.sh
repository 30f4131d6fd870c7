mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_body.py tests/test_gpu_fit.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/pytest_a.log 2>&1
tail -c 300 gpurun_out/pytest_a.log
for sm in 0 1; do
LEMO_SKIN_SMALL=$sm timeout 900 python bench.py --skip-cpu-baseline --skip-extra --skip-infill --skip-prox --skip-perframe --min-seconds 1.0 > gpurun_out/bench_sm$sm.json 2> gpurun_out/bench_sm$sm.err; tail -2 gpurun_out/bench_sm$sm.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_sm$sm.json').read().strip().splitlines()[-1])
print('small=$sm value',d['value'],'ms_per_step',d['ms_per_step'],d['clocks']['sm_mhz'])
PY
done
