mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log 2>&1
tail -c 1500 gpurun_out/pytest_gpu.log
for m in simt tc tc64; do LEMO_VPOSER=$m timeout 120 python tools/diag_vposer_modes.py; done > gpurun_out/diag_vposer.log 2>&1
cat gpurun_out/diag_vposer.log
# compute-sanitizer: memcheck + racecheck over the kernels added this round (small shapes; the tools slow kernels down 10-100x)
SAN="tests/test_gpu_prox_loss.py::test_fused_window_is_bitwise_reproducible tests/test_gpu_fit.py::test_perframe_persistent_kernel_vs_graph_path tests/test_gpu_fit.py::test_infill_pool_equals_single_stage"
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $SAN -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/sanitizer_memcheck.log 2>&1
tail -8 gpurun_out/sanitizer_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 1 python -m pytest tests/test_gpu_prox_loss.py::test_fused_window_is_bitwise_reproducible tests/test_gpu_fit.py::test_perframe_tracks_oracle -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/sanitizer_racecheck.log 2>&1
tail -8 gpurun_out/sanitizer_racecheck.log
