mkdir -p gpurun_out
( LEMO_DEBUG_CHECK=1 timeout 900 python -m pytest tests/test_gpu_prox_loss.py -m gpu -q 2>&1 | tail -80 ) > gpurun_out/pytest_c.log 2>&1
( timeout 1200 python -m pytest tests/test_gpu_body.py tests/test_gpu_prox_loss.py tests/test_gpu_fit.py tests/test_gpu_prox.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/pytest_a.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_stages.csv python tools/run_stage.py prox > gpurun_out/stages_under_ncu.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1200 gpurun_out/pytest_c.log; tail -c 1200 gpurun_out/pytest_a.log; tail -3 gpurun_out/bench_n1.err
