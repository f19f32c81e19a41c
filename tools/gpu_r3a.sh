#!/bin/bash
# joints: GPU parity of the gear / chains scenes, the joint bench (levels vs one thread), the long target-joint unit test, compute-sanitizer incl. joint scenes
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_more_joints.py tests/test_snapshot.py -m gpu -x -q ) 2>&1 | tail -4
timeout 600 python tools/bench_joints.py 100000 30 2>&1 | tail -3 | tee gpurun_out/r3a_bench_joints.txt
timeout 300 python tools/bench_joints.py 1000 30 2>&1 | tail -3 | tee -a gpurun_out/r3a_bench_joints.txt
timeout 900 python tools/run_unittests_cuda.py --timeout 600 --filter 'World_Longer.TargetJointWontCauseTunnelling' 2>&1 | tail -6
SAN_TIMEOUT=600 bash tools/sanitize_gpu.sh 2>&1 | tail -30
