#!/bin/bash
# pre-solve events on the GPU + the listener / destroy unit tests of the reference on CudaWorld
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_events.py tests/test_cudaworld_host.py -m gpu -x -q ) 2>&1 | tail -3
timeout 600 python tools/run_unittests_cuda.py --filter 'World.Listener*:World.CreateDestroy*:World.*Contact*:WorldBody.*' --timeout 60 2>&1 | head -20 | cut -c1-200
