mkdir -p gpurun_out
for pad in 2 1 2 1; do
LEMO_PLANE_PAD=$pad timeout 900 python bench.py --skip-cpu-baseline --skip-extra --skip-infill --skip-prox --skip-perframe --min-seconds 1.0 > gpurun_out/bench_pad$pad.json 2> gpurun_out/bench_pad$pad.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pad$pad.json').read().strip().splitlines()[-1])
print('pad=$pad value',d['value'],'ms_per_step',d['ms_per_step'],'conv_ms',d['roofline']['kernel_ms'],d['clocks']['sm_mhz'])
PY
done
