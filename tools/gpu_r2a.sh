#!/bin/bash
# round 2, first GPU pass: the parity suite (new parallel-TOI and full-size cases) and a baseline bench with the TOI batch anatomy
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2a_pytest.log 2>&1
tail -30 gpurun_out/r2a_pytest.log
PR_TOI_DEBUG=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2a_bench.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['stage_ms_per_step'])"
grep 'toi batch' gpurun_out/r2a_bench.err | tail -6
