#!/bin/bash
# structural mutators on the GPU: the parity test, the whole -m gpu suite, then the reference's unmodified World-API unit tests on CudaWorld
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2o_pytest_gpu.log 2>&1
tail -16 gpurun_out/r2o_pytest_gpu.log
timeout 1500 python tools/run_unittests_cuda.py --timeout 60 > gpurun_out/r2o_unittests_cuda.txt 2> gpurun_out/r2o_unittests_cuda.err
head -60 gpurun_out/r2o_unittests_cuda.txt | cut -c1-230
