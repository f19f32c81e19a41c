#!/bin/bash
# ncu --set full on k_block_solve (tumbler arena, 1024 worlds) with source-line attribution
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_block_solve$' -c 1 \
  -o /tmp/prof_b python bench.py --workload tumblers --worlds 1024 --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/prof_b.log 2>&1
ncu -i /tmp/prof_b.ncu-rep --page raw --csv > gpurun_out/prof_b_raw.csv 2>/dev/null
ncu -i /tmp/prof_b.ncu-rep --page source --csv --print-source cuda > gpurun_out/prof_b_source.csv 2>gpurun_out/prof_b_source.err
ncu -i /tmp/prof_b.ncu-rep --page details > gpurun_out/prof_b_details.txt 2>/dev/null
ls -la gpurun_out/prof_b* /tmp/prof_b.ncu-rep
python profiles/summarize_full.py gpurun_out/prof_b_raw.csv | cut -c1-250
