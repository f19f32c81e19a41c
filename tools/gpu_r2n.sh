#!/bin/bash
# tumbler arena: block solver (default up to 2560 worlds) vs arena-wide dataflow solver (debug bit 5 = 32) vs per-colour launches (32+128)
mkdir -p gpurun_out
run() { # worlds flags
  timeout 400 python bench.py --workload tumblers --worlds $1 --debug-flags $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_$1_$2.json 2> gpurun_out/r2n_$1_$2.err
  python -c "
import json;d=json.load(open('gpurun_out/r2n_$1_$2.json'));s=d['stage_ms_per_step'];print('$1 worlds flags $2:', round(d['ms_per_step'],3), round(d['value']/1e6,1), 'M/s solve', round(s['velocity']+s['position'],3), 'setup', round(s['setup'],3), 'color', round(s['color'],3), d['gpu_launches'])" || tail -5 gpurun_out/r2n_$1_$2.err
}
run 1024 0; run 1024 32; run 8192 0; run 8192 160
