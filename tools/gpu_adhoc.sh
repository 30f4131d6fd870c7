mkdir -p gpurun_out
for mode in wt pair; do
LEMO_CONV=$mode timeout 900 python bench.py --skip-cpu-baseline --skip-extra --skip-infill --skip-perframe --min-seconds 0.3 > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; tail -2 gpurun_out/bench_$mode.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$mode.json').read().strip().splitlines()[-1])
print('$mode value',d['value'],'ms_per_step',d['ms_per_step'],'prox',{k:v for k,v in (d.get('prox') or {}).items() if 'ms_per' in k})
PY
done
