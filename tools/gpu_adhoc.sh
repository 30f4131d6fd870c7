mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_scripts_surface.py tests/test_gpu_body.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/pytest_a.log 2>&1
tail -c 2500 gpurun_out/pytest_a.log
( timeout 900 python -m pytest tests/test_gpu_loops_baseline.py -m gpu -q -k perframe 2>&1 | tail -40 ) > gpurun_out/pytest_b.log 2>&1
tail -c 1500 gpurun_out/pytest_b.log
timeout 1200 python bench.py --skip-cpu-baseline --skip-prox > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
