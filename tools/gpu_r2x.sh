#!/bin/bash
# randomized-linking union-find: parity + islands stage time (baseline 0.33 ms)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_batch.py tests/test_destroy_body_shape.py -m gpu -x -q ) 2>&1 | tail -5
timeout 600 python bench.py --steps 24 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
tail -3 gpurun_out/r2x_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[0]);print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'])"
