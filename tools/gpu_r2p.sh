#!/bin/bash
# A/B: colouring rounds in one block (debug bit 8 = launches), ECL-CC style island init (bit 9 = unions alone)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_batch.py -m gpu -x -q ) 2>&1 | tail -3
for f in 0 256 512; do
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --debug-flags $f > gpurun_out/r2p_$f.json 2> gpurun_out/r2p_$f.err
python -c "
import json;d=json.load(open('gpurun_out/r2p_$f.json'));print($f, d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'])"
done
