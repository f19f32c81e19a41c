#!/bin/bash
# one `ncu --set full` capture of the regular-phase narrowphase launch with source-line attribution
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_update_contacts$' -s 1 -c 1 \
  -o /tmp/prof_n python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/prof_n.log 2>&1
ncu -i /tmp/prof_n.ncu-rep --page raw --csv > gpurun_out/prof_n_raw.csv 2>/dev/null
ncu -i /tmp/prof_n.ncu-rep --page source --csv --print-source cuda > gpurun_out/prof_n_source_cuda.csv 2>gpurun_out/prof_n_source.err
ls -la gpurun_out /tmp/prof_n.ncu-rep
python profiles/summarize_full.py gpurun_out/prof_n_raw.csv
head -c 1500 gpurun_out/prof_n_source_cuda.csv
