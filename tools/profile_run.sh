#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench step + full ncu capture of the dominant kernel.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 30 -c 2 -o gpurun_out/prof_conv -f \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
