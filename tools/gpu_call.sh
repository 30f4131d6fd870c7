#!/bin/bash
# One GPU lease (run under gpurun): ./tools/gpu_call.sh [tests] [bench] [ref] [launches] [ncu_conv] [ncu_lbs] [stages]
# Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by tools/summarize_profiles.py.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
FAST="--steps 2 --warmup 1 --no-graph --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill"
for step in "$@"; do
  case $step in
    tests)    ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log ;;
    tests_all) ( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log ;;
    bench)    timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err ;;
    ref)      timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
                python bench.py $FAST > gpurun_out/bench_under_ncu.log 2>&1 ;;
    ncu_conv) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc" -s 40 -c 2 -o gpurun_out/prof_conv_tc -f \
                python bench.py $FAST > gpurun_out/bench_under_ncu2.log 2>&1 ;;
    ncu_lbs)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_blend_v2|k_skin_tc|k_pose_chain_fwd|k_joints_fwd" -s 4 -c 4 \
                -o gpurun_out/prof_lbs -f python tools/diag_lbs.py 120 > gpurun_out/diag_under_ncu.log 2>&1
              timeout 120 python tools/diag_lbs.py 120 300 > gpurun_out/diag_lbs.log 2>&1 ;;
    stages)   # launch lists of the secondary stages (infill pre-stage, per-frame, PROX window) for planning
              timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_stages.csv \
                python tools/run_stage.py prox perframe > gpurun_out/stages_under_ncu.log 2>&1; tail -3 gpurun_out/stages_under_ncu.log ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la gpurun_out | tail -12
