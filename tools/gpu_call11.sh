#!/bin/bash
# full ncu capture of the utility SGEMM launches of one fitting iteration (VPoser MLP fwd/bwd, LBS adjoint contraction)
mkdir -p gpurun_out
FAST="--steps 2 --warmup 1 --no-graph --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm" -s 31 -c 7 -o gpurun_out/prof_gemm -f \
    python bench.py $FAST > gpurun_out/bench_under_ncu3.log 2>&1
tail -3 gpurun_out/bench_under_ncu3.log
