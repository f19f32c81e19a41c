#!/bin/bash
# every kernel launch of 2 timed steps of the 1M mixed scene with its device time (cold-cache, serialised under ncu:
# compare SHARES with bench.py's stage times, not absolutes)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_mixed.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cuda-profiler > gpurun_out/launches_mixed.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_mixed.csv > gpurun_out/launches_mixed.txt
head -70 gpurun_out/launches_mixed.txt
