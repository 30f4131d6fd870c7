#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_body.py -x -q > gpurun_out/pytest_body.log 2>&1; tail -15 gpurun_out/pytest_body.log
timeout 120 python tools/diag_lbs.py 120 300 2>&1 | tee gpurun_out/diag_lbs.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_blend_tf32|k_skin|k_chain_fwd|k_pose_to_rot|k_joints" -c 24 python tools/diag_lbs.py 120 2>&1 | grep -v "^==PROF" | grep -i "k_\|duration" | paste - - | awk '{print $1, $NF}' | tail -24
