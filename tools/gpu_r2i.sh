#!/bin/bash
mkdir -p gpurun_out
for lib in libplayrho_b200_toitiming.so libplayrho_b200_tt_nocache.so libplayrho_b200_tt_generic_nocache.so; do
PLAYRHO_B200_LIB=$PWD/playrho_b200/$lib timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2i.json 2> gpurun_out/r2i.err
echo $lib; grep -E 'late_timing|toi_timing' gpurun_out/r2i.err | tail -2
python -c "
import json;d=json.load(open('gpurun_out/r2i.json'));print(d['ms_per_step'],d['stage_ms_per_step']['toi'],d['stage_ms_per_step']['position'])"
done
