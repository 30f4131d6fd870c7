#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_conv_wt.py --experiments > gpurun_out/diag_wt.log 2>&1
grep -v "Warn\|detach\|tag, q_rel" gpurun_out/diag_wt.log | tail -24
for m in 0 1 8192; do timeout 120 python tools/diag_conv_wt.py $m 2>&1 | grep -v "Warn\|detach\|tag, q_rel" | tail -3; done
echo "== simt then pair in one process"
timeout 120 python tools/diag_conv_wt.py 0 1 2>&1 | grep -v "Warn\|detach\|tag, q_rel" | tail -4
echo "== sanitizer"
CUDA_LAUNCH_BLOCKING=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/diag_conv_wt.py 0 1 > gpurun_out/sanitizer.log 2>&1
grep -v "Warn\|detach\|tag, q_rel" gpurun_out/sanitizer.log | head -60
