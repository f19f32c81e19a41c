#!/bin/bash
# ncu --set full captures of the step's main kernels on the 1M mixed scene (one timed step after 300 settle
# steps), reduced ON the box to CSV pages (the .ncu-rep files are too large for gpurun_out's 64 MiB).
mkdir -p gpurun_out
A='^(k_update_contacts|k_bvh_query_coop|k_bvh_query|k_tb_solve|k_isl_union|k_isl_flatten|k_isl_prepare|k_build_constraints|k_toi_compute|k_sleep_and_stats|k_sync_proxies|k_toi_claim|k_create_contacts|k_writeback_bodies|k_mark_constraints|k_destroy_test|k_tb_query|k_tb_add)$'
run() { # name regex count
  ncu --set full --clock-control none --profile-from-start off -k regex:"$2" -c $3 \
    -o /tmp/prof_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/prof_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
}
run a "$A" ${NA:-34}
run v '^k_solve_velocity$' 3
run p '^k_solve_position$' 2
python profiles/summarize_full.py gpurun_out/prof_a_raw.csv gpurun_out/prof_v_raw.csv gpurun_out/prof_p_raw.csv > gpurun_out/full_summary.txt
cat gpurun_out/full_summary.txt | cut -c1-220
