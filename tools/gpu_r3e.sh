#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_batch.py tests/test_events.py -m gpu -x -q ) 2>&1 | tail -3
PR_SYNC_DEBUG=1 timeout 600 python bench.py --steps 24 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err
grep "host syncs" gpurun_out/r3e_bench.err | sed -n '320,322p'
python -c "
import json;d=json.loads(open('gpurun_out/r3e_bench.json').read().strip().splitlines()[0]);print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['e2e']['ms_per_step'])"
timeout 600 python bench.py --steps 24 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[0]);print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['e2e']['ms_per_step'])"
