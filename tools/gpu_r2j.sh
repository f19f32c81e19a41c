#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma" ) 2>&1 | tail -3
for f in 0 64; do
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --debug-flags $f > gpurun_out/r2j_$f.json 2> gpurun_out/r2j_$f.err
python -c "
import json;d=json.load(open('gpurun_out/r2j_$f.json'));print($f, d['ms_per_step'],d['stage_ms_per_step']['narrow'],d['roofline']['frac'])"
done
