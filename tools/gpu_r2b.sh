#!/bin/bash
# TOI anatomy: the -DPR_TOI_TIMING build (clock64 deltas of the event solve and of the single-block late rounds)
mkdir -p gpurun_out
PLAYRHO_B200_LIB=$PWD/playrho_b200/libplayrho_b200_toitiming.so PR_TOI_DEBUG=1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
grep -E 'late_timing|toi_timing' gpurun_out/r2b_bench.err | tail -8
python -c "
import json;d=json.load(open('gpurun_out/r2b_bench.json'));print(d['ms_per_step'],d['stage_ms_per_step'])"
