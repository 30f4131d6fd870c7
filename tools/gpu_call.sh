#!/bin/bash
# One GPU lease (run under gpurun): ./tools/gpu_call.sh [smoke] [tests] [tests_all] [bench] [ref] [launches] [ncu_conv] [ncu_lbs] [ncu_stages] [stages]
#                                                       [memcheck] [racecheck] [bench2]
# Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by tools/summarize_profiles.py.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
FAST="--steps 2 --warmup 1 --no-graph --skip-cpu-baseline --skip-prox --skip-perframe --skip-infill --skip-extra"
SAN="tests/test_gpu_prox_loss.py::test_fused_window_is_bitwise_reproducible tests/test_gpu_fit.py::test_perframe_persistent_kernel_vs_graph_path tests/test_gpu_fit.py::test_infill_pool_equals_single_stage"
for step in "$@"; do
  case $step in
    smoke)    ( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log ;;
    tests)    ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log ;;
    tests_all) ( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log ;;
    bench)    timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err ;;
    bench2)   timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 \
                --skip-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err ;;
    ref)      timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
                python bench.py $FAST > gpurun_out/bench_under_ncu.log 2>&1 ;;
    ncu_conv) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc" -s 40 -c 2 -o gpurun_out/prof_conv_tc -f \
                python bench.py $FAST > gpurun_out/bench_under_ncu2.log 2>&1 ;;
    ncu_lbs)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_blend_v2|k_skin_tc|k_pose_chain_fwd|k_joints_fwd" -s 4 -c 4 \
                -o gpurun_out/prof_lbs -f python tools/diag_lbs.py 120 > gpurun_out/diag_under_ncu.log 2>&1
              timeout 120 python tools/diag_lbs.py 120 300 > gpurun_out/diag_lbs.log 2>&1 ;;
    ncu_stages) # full captures of the round-2 kernels: persistent per-frame kernel, PROX window kernels, AE fine-tune kernels
              timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_perframe_mega" -c 1 -o gpurun_out/prof_perframe -f \
                python tools/run_stage.py perframe > gpurun_out/ncu_pf.log 2>&1
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scene_query|k_skin_bwd|k_dx_tallk|k_prox_scene|k_prox_rows|k_chain_bwd" -c 8 \
                -o gpurun_out/prof_prox -f python tools/run_stage.py prox > gpurun_out/ncu_prox.log 2>&1
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wgrad|k_conv3x3" -s 200 -c 6 -o gpurun_out/prof_ae -f \
                python tools/run_stage.py infill > gpurun_out/ncu_ae.log 2>&1 ;;
    stages)   # launch lists of the secondary stages (per-frame, PROX window, infill) -- one file each
              for s in prox perframe infill; do
                timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$s.csv \
                  python tools/run_stage.py $s > gpurun_out/stage_${s}_under_ncu.log 2>&1
              done ;;
    memcheck) ( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $SAN -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/sanitizer_memcheck.log 2>&1
              tail -4 gpurun_out/sanitizer_memcheck.log ;;
    racecheck) for s in perframe prox infill; do
                timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/run_stage.py $s > gpurun_out/sanitizer_racecheck_$s.log 2>&1
                tail -3 gpurun_out/sanitizer_racecheck_$s.log
              done ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la gpurun_out | tail -12
