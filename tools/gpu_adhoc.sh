mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -80 ) > gpurun_out/pytest_gpu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_stages.csv python tools/run_stage.py prox infill > gpurun_out/stages_under_ncu.log 2>&1
timeout 1200 python bench.py --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 2500 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_n1.err
