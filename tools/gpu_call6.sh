#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_conv_wt.py --experiments > gpurun_out/diag_wt.log 2>&1
grep -v "Warn\|detach\|tag, q_rel" gpurun_out/diag_wt.log | tail -34
for m in 8192 8256; do timeout 120 python tools/diag_conv_wt.py $m 2>&1 | grep -v "Warn\|detach\|tag, q_rel" | tail -3; done
FAST="--steps 20 --warmup 5 --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill"
for cfg in "wt"; do
  echo "== LEMO_CONV=$cfg"
  LEMO_CONV=$cfg timeout 300 python bench.py $FAST 2> gpurun_out/bench_$cfg.err | tee gpurun_out/bench_$cfg.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline_lbs']['ms'])
"
done
