mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'^(k_update_contacts|k_bvh_query|k_toi_solve|k_isl_union|k_isl_flatten|k_isl_prepare|k_build_constraints|k_toi_compute|k_sleep_and_stats|k_sync_proxies|k_toi_claim|k_toi_select|k_create_contacts)$' -c 45 \
  -o gpurun_out/prof_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/prof_a.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'^(k_solve_velocity|k_solve_position|k_warm_start|k_isl_vel_check)$' -c 18 \
  -o gpurun_out/prof_b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/prof_b.log 2>&1
ls -la gpurun_out
