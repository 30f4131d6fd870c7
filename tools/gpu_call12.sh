#!/bin/bash
# parity suite + headline bench line (no secondaries) with and without the cluster split-K GEMM
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
FASTB="--steps 30 --warmup 5 --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill"
for c in 1 0; do
  LEMO_GEMM_CLUSTER=$c timeout 300 python bench.py $FASTB > gpurun_out/bench_gc$c.json 2> gpurun_out/bench_gc$c.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_gc$c.json').read().strip().splitlines()[-1])
print('cluster=$c', 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'lbs', d['roofline_lbs']['ms'])"
done
