#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_conv_wt.py --experiments > gpurun_out/diag_wt.log 2>&1
cat gpurun_out/diag_wt.log | tail -20
( timeout 600 python -m pytest tests/test_gpu_priors.py tests/test_gpu_fit.py -m gpu -q ) > gpurun_out/pytest_wt.log 2>&1
tail -15 gpurun_out/pytest_wt.log
FAST="--steps 20 --warmup 5 --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill"
for cfg in "pair 64" "pair 0" "wt 0"; do
  set -- $cfg
  echo "== LEMO_CONV=$1 LEMO_GEMM_BM=$2"
  LEMO_CONV=$1 LEMO_GEMM_BM=$2 timeout 300 python bench.py $FAST 2> gpurun_out/bench_$1_$2.err | tee gpurun_out/bench_$1_$2.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline_lbs']['ms'])
"
done
