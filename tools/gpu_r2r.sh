#!/bin/bash
# packed child records + TOI group hint: parity subset and the bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_destroy_body_shape.py tests/test_snapshot.py -m gpu -x -q ) 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2r.json 2> gpurun_out/r2r.err
python -c "
import json;d=json.load(open('gpurun_out/r2r.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['roofline']['frac'])"
