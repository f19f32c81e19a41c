#!/bin/bash
mkdir -p gpurun_out
for T in 64 128; do
PLAYRHO_B200_LIB=$PWD/playrho_b200/libplayrho_b200_toitiming.so PR_BLK_THREADS=$T timeout 400 python bench.py --workload tumblers --worlds 1024 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2e.json 2> gpurun_out/r2e.err
grep blk_timing gpurun_out/r2e.err | tail -1
python -c "
import json;d=json.load(open('gpurun_out/r2e.json'));print($T, round(d['ms_per_step'],3), round(d['stage_ms_per_step']['velocity'],3))"
done
