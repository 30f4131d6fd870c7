#!/bin/bash
# One GPU call: parity tests, bench line, reference arm, ncu launch list, full captures of the dominant conv kernel and the LBS kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
FAST="--steps 2 --warmup 1 --no-graph --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py $FAST > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc" -s 40 -c 2 -o gpurun_out/prof_conv_tc -f \
    python bench.py $FAST > gpurun_out/bench_under_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_blend_v2|k_skin_tc|k_pose_chain_fwd|k_joints_fwd" -s 4 -c 4 -o gpurun_out/prof_lbs -f \
    python tools/diag_lbs.py 120 > gpurun_out/diag_under_ncu.log 2>&1
timeout 120 python tools/diag_lbs.py 120 300 > gpurun_out/diag_lbs.log 2>&1
ls -la gpurun_out | tail -12
