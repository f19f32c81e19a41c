#!/bin/bash
# warp-cooperative TOI event solve, variants of the two serial passes (inline / out of line, UnitVec cache on / off)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_events.py tests/test_sub_stepping.py -m gpu -x -q ) 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2l.json 2> gpurun_out/r2l.err
python -c "
import json;d=json.load(open('gpurun_out/r2l.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
for v in toitiming tt_cache tt_call; do
PLAYRHO_B200_LIB=$PWD/playrho_b200/libplayrho_b200_$v.so PR_TOI_DEBUG=1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2l_$v.json 2> gpurun_out/r2l_$v.err
echo $v; grep -E 'late_timing|toi_timing' gpurun_out/r2l_$v.err | tail -2
python -c "
import json;d=json.load(open('gpurun_out/r2l_$v.json'));print(d['ms_per_step'],d['stage_ms_per_step']['toi'])"
done
