#!/bin/bash
# LBVH: boxes re-fitted over the last rebuild's topology between rebuilds (PR_BVH_REFIT_STEPS = 0 / 1 / 3 / 7); joints on the GPU
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_distance_joint.py tests/test_bulk_velocity.py tests/test_batch.py -m gpu -x -q ) 2>&1 | tail -3
for k in 0 1 3 7; do
PR_BVH_REFIT_STEPS=$k timeout 400 python bench.py --steps 24 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2v_$k.json 2> gpurun_out/r2v_$k.err
python -c "
import json;d=json.load(open('gpurun_out/r2v_$k.json'));print($k, d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'])"
done
