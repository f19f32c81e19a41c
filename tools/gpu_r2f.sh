#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_batch.py tests/test_gpu_parity.py -m gpu -x -q -k "batch or tumbler" ) 2>&1 | tail -3
run() { # worlds flags threads
  PR_BLK_THREADS=$3 timeout 400 python bench.py --workload tumblers --worlds $1 --debug-flags $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_$1_$2_$3.json 2> gpurun_out/r2f_$1_$2_$3.err
  python -c "
import json;d=json.load(open('gpurun_out/r2f_$1_$2_$3.json'));s=d['stage_ms_per_step'];print('$1 worlds flags $2 threads $3:', round(d['ms_per_step'],3), 'solve', round(s['velocity']+s['position'],3), 'setup', round(s['setup'],3), 'color', round(s['color'],3))" || tail -5 gpurun_out/r2f_$1_$2_$3.err
}
run 512 0 32; run 1024 0 32; run 1024 0 64; run 1024 0 128; run 2048 0 32; run 8192 0 32; run 8192 0 64
PLAYRHO_B200_LIB=$PWD/playrho_b200/libplayrho_b200_toitiming.so PR_BLK_THREADS=32 timeout 400 python bench.py --workload tumblers --worlds 1024 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2e.json 2> gpurun_out/r2e.err
grep blk_timing gpurun_out/r2e.err | tail -1
