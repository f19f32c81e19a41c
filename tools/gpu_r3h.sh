#!/bin/bash
mkdir -p gpurun_out
PR_SYNC_DEBUG=1 timeout 600 python bench.py --steps 24 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err
grep "host syncs" gpurun_out/r3h_bench.err | sed -n '320,321p'
python -c "
import json;d=json.loads(open('gpurun_out/r3h_bench.json').read().strip().splitlines()[0]);print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['e2e']['ms_per_step'])"
