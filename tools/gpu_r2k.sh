#!/bin/bash
# warp-cooperative TOI event solve: parity (TOI-heavy GPU tests) + bench + timing-build anatomy
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_events.py tests/test_sub_stepping.py -m gpu -x -q ) 2>&1 | tail -5
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2k.json 2> gpurun_out/r2k.err
python -c "
import json;d=json.load(open('gpurun_out/r2k.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
PLAYRHO_B200_LIB=$PWD/playrho_b200/libplayrho_b200_toitiming.so PR_TOI_DEBUG=1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_tt.json 2> gpurun_out/r2k_tt.err
grep -E 'late_timing|toi_timing' gpurun_out/r2k_tt.err | tail -4
