#!/bin/bash
# LBS forward round: body parity tests, eager/graph timing per back end, per-kernel launch times, full ncu capture of the two GEMM kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_body.py -x -q > gpurun_out/pytest_body.log 2>&1; tail -15 gpurun_out/pytest_body.log
timeout 120 python tools/diag_lbs.py 120 300 2>&1 | tee gpurun_out/diag_lbs.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_blend|k_skin|k_pose_chain|k_joints" -c 24 --csv --log-file gpurun_out/launches_lbs.csv python tools/diag_lbs.py 120 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_lbs.csv') if not l.startswith('=='))]
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:40], r[h.index('Metric Value')], r[h.index('Metric Unit')])
PY
if [ "$1" = "full" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_blend_v2|k_skin_tc" -s 4 -c 4 -o gpurun_out/prof_lbs -f python tools/diag_lbs.py 120 > gpurun_out/diag_under_ncu.log 2>&1
fi
