#!/bin/bash
# Run on the GPU box (under gpurun): launch list of bench steps + full ncu captures of the dominant kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu-baseline --skip-prox --skip-perframe > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc" -s 20 -c 2 -o gpurun_out/prof_conv_tc -f \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu-baseline --skip-prox --skip-perframe > gpurun_out/bench_under_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gemm|k_skin_bwd|k_tc_last_bwd" -s 30 -c 12 -o gpurun_out/prof_misc -f \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu-baseline --skip-prox --skip-perframe > gpurun_out/bench_under_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_blend_tf32|k_skin_fwd" -s 4 -c 4 -o gpurun_out/prof_lbs -f \
    python tools/diag_blend.py > gpurun_out/diag_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
