#!/bin/bash
# the whole -m gpu suite with durations
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/pytest_gpu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log
