#!/bin/bash
# round 2: launch list (2 timed steps) + ncu --set full of the step's main kernels on the 1M mixed scene, reduced ON the box to CSV
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --cuda-profiler > gpurun_out/r02_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches.txt
head -45 gpurun_out/r02_launches.txt
A='^(k_update_contacts|k_bvh_query_coop|k_bvh_query|k_tb_solve|k_tb_late|k_isl_union|k_isl_flatten|k_build_constraints|k_toi_compute|k_sync_proxies|k_color_claim|k_color_commit|k_flow_words|k_destroy_test|k_mark_constraints|k_bvh_refit|k_create_contacts)$'
run() { # name regex count
  timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"$2" -c $3 \
    -o /tmp/prof_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --cuda-profiler > gpurun_out/r02_prof_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r02_prof_$1_raw.csv 2>/dev/null
}
run a "$A" ${NA:-40}
run f 'k_flow_pass' 6
python profiles/summarize_full.py gpurun_out/r02_prof_a_raw.csv gpurun_out/r02_prof_f_raw.csv > gpurun_out/r02_full_summary.txt
cut -c1-250 gpurun_out/r02_full_summary.txt
