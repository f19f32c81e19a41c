#!/bin/bash
# the new joint types on the GPU + the World-API unit tests
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_more_joints.py tests/test_distance_joint.py tests/test_friction_motor_joint.py tests/test_gpu_parity.py -m gpu -x -q ) 2>&1 | tail -5
timeout 1500 python tools/run_unittests_cuda.py --timeout 60 > gpurun_out/r2y_unittests_cuda.txt 2> gpurun_out/r2y_unittests_cuda.err
head -32 gpurun_out/r2y_unittests_cuda.txt | cut -c1-220
