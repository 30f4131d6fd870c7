mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log 2>&1
tail -c 1200 gpurun_out/pytest_gpu.log
SAN="tests/test_gpu_prox_loss.py::test_fused_window_is_bitwise_reproducible tests/test_gpu_fit.py::test_perframe_persistent_kernel_vs_graph_path tests/test_gpu_fit.py::test_infill_pool_equals_single_stage"
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $SAN -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/sanitizer_memcheck.log 2>&1
tail -6 gpurun_out/sanitizer_memcheck.log
( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 1 python -m pytest tests/test_gpu_prox_loss.py::test_fused_window_is_bitwise_reproducible tests/test_gpu_fit.py::test_perframe_tracks_oracle -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/sanitizer_racecheck.log 2>&1
tail -6 gpurun_out/sanitizer_racecheck.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_stages.csv python tools/run_stage.py prox > gpurun_out/stages_under_ncu.log 2>&1
timeout 1200 python bench.py --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
