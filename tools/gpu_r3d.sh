#!/bin/bash
mkdir -p gpurun_out
PR_SYNC_DEBUG=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err
grep "host syncs" gpurun_out/r3d_bench.err | tail -3
