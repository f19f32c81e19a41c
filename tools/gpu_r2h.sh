#!/bin/bash
mkdir -p gpurun_out
for strip in 4000 8000; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --strip $strip > gpurun_out/r2h_$strip.json 2> gpurun_out/r2h_$strip.err
python -c "
import json;d=json.load(open('gpurun_out/r2h_$strip.json'));print($strip, d['ms_per_step'],d['config']['toi_events_per_step'],d['config']['toi_rounds_per_step'],d['config']['contacts'],d['config']['colors'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
done
