#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/run_unittests_cuda.py --timeout 120 --filter 'WheelJointConf.*:GearJoint*:PrismaticJointConf.*' 2>&1 | head -12 | cut -c1-200
PR_SYNC_DEBUG=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
grep "host syncs" gpurun_out/r3c_bench.err | tail -8
python -c "
import json;d=json.loads(open('gpurun_out/r3c_bench.json').read().strip().splitlines()[0]);print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'])"
