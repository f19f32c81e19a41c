#!/bin/bash
# hybrid colouring rounds (3 launched rounds + one block for the tail; debug bit 8 = launches only); k_flow_pass occupancy variants
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_batch.py tests/test_destroy_body_shape.py -m gpu -x -q ) 2>&1 | tail -3
for f in 0 256; do
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --debug-flags $f > gpurun_out/r2q_$f.json 2> gpurun_out/r2q_$f.err
python -c "
import json;d=json.load(open('gpurun_out/r2q_$f.json'));print($f, d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'])"
done
for v in flow4 flow2; do
PLAYRHO_B200_LIB=$PWD/playrho_b200/libplayrho_b200_$v.so timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2q_$v.json 2> gpurun_out/r2q_$v.err
python -c "
import json;d=json.load(open('gpurun_out/r2q_$v.json'));print('$v', d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
done
