mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_prox_loss.py tests/test_gpu_ae.py tests/test_gpu_infill.py tests/test_scripts_surface.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/pytest_a.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_priors.py -m gpu -q -k "scene or chamfer or aa_outputs or tgm" 2>&1 | tail -30 ) > gpurun_out/pytest_b.log 2>&1
( LEMO_DEBUG_CHECK=1 timeout 600 python -m pytest tests/test_gpu_prox_loss.py -m gpu -q -x -k "fall_back" 2>&1 | tail -60 ) > gpurun_out/pytest_c.log 2>&1
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1500 gpurun_out/pytest_a.log; tail -5 gpurun_out/pytest_b.log; tail -30 gpurun_out/pytest_c.log
