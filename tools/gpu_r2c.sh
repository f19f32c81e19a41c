#!/bin/bash
# block solver (one CTA per tumbler world): parity on the GPU, then the tumbler arena at 8192 and 1024 worlds, variants
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_batch.py tests/test_gpu_parity.py -m gpu -x -q -k "batch or tumbler" ) 2>&1 | tail -3
run() { # worlds flags threads colors
  PR_BLK_THREADS=$3 PR_BLK_COLORS=$4 timeout 400 python bench.py --workload tumblers --worlds $1 --debug-flags $2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_t_$1_$2_$3_$4.json 2> gpurun_out/r2c_t_$1_$2_$3_$4.err
  python -c "
import json;d=json.load(open('gpurun_out/r2c_t_$1_$2_$3_$4.json'));print('$1 worlds flags $2 threads $3 colors $4:', round(d['ms_per_step'],3), d['config'].get('colors'), {k: round(v,3) for k,v in d['stage_ms_per_step'].items() if v > 0.05})" || tail -5 gpurun_out/r2c_t_$1_$2_$3_$4.err
}
run 1024 32 128 16
run 1024 0 128 16
run 1024 0 256 16
run 1024 0 64 16
run 1024 0 128 8
run 1024 0 128 24
run 8192 0 128 16
