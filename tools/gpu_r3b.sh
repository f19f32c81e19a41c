#!/bin/bash
# the extended acceptance suite (all joint unit-test files + Dump + WorldManifold on CudaWorld) and initcheck after the zero-fill change
mkdir -p gpurun_out
timeout 2400 python tools/run_unittests_cuda.py --timeout 120 > gpurun_out/r3b_unittests_cuda.txt 2> gpurun_out/r3b_unittests_cuda.err
head -60 gpurun_out/r3b_unittests_cuda.txt | cut -c1-230
timeout 1200 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize_scene.py 8 > gpurun_out/sanitizer_initcheck.txt 2>&1
echo "== initcheck rc=$?"; grep -E "ERROR SUMMARY|sanitize scene done|skipped" gpurun_out/sanitizer_initcheck.txt | head
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_batch.py -m gpu -x -q ) 2>&1 | tail -3
