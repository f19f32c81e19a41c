#!/bin/bash
# final build of the round: whole -m gpu suite, acceptance unit tests, default bench + reference arm, launch list + ncu summaries
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r3f_pytest_gpu.log 2>&1
tail -6 gpurun_out/r3f_pytest_gpu.log
timeout 2400 python tools/run_unittests_cuda.py --timeout 120 > gpurun_out/r3f_unittests_cuda.txt 2> gpurun_out/r3f_unittests_cuda.err
head -3 gpurun_out/r3f_unittests_cuda.txt | cut -c1-200; grep -A12 "not passing" gpurun_out/r3f_unittests_cuda.txt | cut -c1-160
( time timeout 900 python bench.py ) > gpurun_out/r3f_bench_default.json 2> gpurun_out/r3f_bench_default.err
tail -3 gpurun_out/r3f_bench_default.err
python -c "
import json;d=json.loads(open('gpurun_out/r3f_bench_default.json').read().strip().splitlines()[0]);print(d['ms_per_step'],d['e2e']['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['roofline']['frac'], d.get('tumblers_n1',{}).get('ms_per_step'), d['cpu_baseline'])"
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r3f_bench_reference.json 2> gpurun_out/r3f_bench_reference.err
cut -c1-200 gpurun_out/r3f_bench_reference.json
NA=40 bash profiles/capture_r02.sh > gpurun_out/r3f_capture.log 2>&1
tail -5 gpurun_out/r3f_capture.log | cut -c1-200
