#!/bin/bash
# ncu --set full captures of the step's main kernels on the 1M mixed scene (one timed step after 300 settle
# steps), reduced ON the box to CSV pages (the .ncu-rep files are too large for gpurun_out's 64 MiB).
mkdir -p gpurun_out
A='^(k_update_contacts|k_bvh_query|k_toi_solve|k_isl_union|k_isl_flatten|k_isl_prepare|k_build_constraints|k_toi_compute|k_sleep_and_stats|k_sync_proxies|k_toi_claim|k_create_contacts|k_writeback_bodies|k_mark_constraints)$'
B='^(k_solve_velocity|k_solve_position|k_warm_start|k_isl_vel_check)$'
ncu --set full --clock-control none --profile-from-start off -k regex:"$A" -c ${NA:-30} \
  -o /tmp/prof_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/prof_a.log 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:"$B" -c ${NB:-6} \
  -o /tmp/prof_b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/prof_b.log 2>&1
for p in a b; do
  ncu -i /tmp/prof_$p.ncu-rep --page raw --csv > gpurun_out/prof_${p}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$p.ncu-rep --page source --csv 2>/dev/null | head -c 24000000 > gpurun_out/prof_${p}_source.csv
  ls -la /tmp/prof_$p.ncu-rep
done
gzip -f gpurun_out/prof_*_source.csv
ls -la gpurun_out
