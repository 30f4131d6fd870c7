#!/bin/bash
# First GPU call of the next round: measure the two experimental paths written after round 1's GPU budget was spent.
mkdir -p gpurun_out
for m in simt tc tc64; do LEMO_VPOSER=$m timeout 120 python tools/diag_vposer_modes.py; done 2>&1 | grep -v Warning | tee gpurun_out/diag_vposer_modes.log
FASTB="--steps 30 --warmup 5 --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill"
for m in simt tc64; do
  LEMO_VPOSER=$m timeout 300 python bench.py $FASTB > gpurun_out/bench_vp_$m.json 2> gpurun_out/bench_vp_$m.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_vp_$m.json').read().strip().splitlines()[-1])
print('LEMO_VPOSER=$m', 'value', d['value'], 'ms/step', d['ms_per_step'])"
done
LEMO_VPOSER=tc64 timeout 600 python -m pytest tests/test_gpu_priors.py tests/test_gpu_fit.py -x -q 2>&1 | tail -5
