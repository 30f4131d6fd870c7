mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_body.py tests/test_gpu_prox_loss.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/pytest_a.log 2>&1
tail -c 600 gpurun_out/pytest_a.log
timeout 900 python bench.py --skip-cpu-baseline --skip-extra --skip-infill --skip-perframe --min-seconds 0.3 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err; tail -2 gpurun_out/bench_pf.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pf.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'],'prox',{k:v for k,v in (d.get('prox') or {}).items() if 'ms_per' in k})
PY
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/run_stage.py perframe > gpurun_out/sanitizer_racecheck_perframe.log 2>&1; tail -5 gpurun_out/sanitizer_racecheck_perframe.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/run_stage.py prox > gpurun_out/sanitizer_racecheck_prox.log 2>&1; tail -5 gpurun_out/sanitizer_racecheck_prox.log
