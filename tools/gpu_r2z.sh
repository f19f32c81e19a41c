#!/bin/bash
# final run of the round: whole -m gpu suite, the World-API unit tests on CudaWorld, default bench (as the driver runs it), reference arm, ncu capture
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2z_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2z_pytest_gpu.log
timeout 1500 python tools/run_unittests_cuda.py --timeout 60 > gpurun_out/r2z_unittests_cuda.txt 2> gpurun_out/r2z_unittests_cuda.err
head -30 gpurun_out/r2z_unittests_cuda.txt | cut -c1-200
( time timeout 900 python bench.py ) > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err
tail -4 gpurun_out/r2z_bench_default.err
python -c "
import json;d=json.loads(open('gpurun_out/r2z_bench_default.json').read().strip().splitlines()[0]);print(d['ms_per_step'],d['e2e']['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['roofline']['frac'], d.get('tumblers_n1',{}).get('ms_per_step'), d['cpu_baseline'])"
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
cut -c1-300 gpurun_out/r2z_bench_reference.json
bash profiles/capture_r02.sh > gpurun_out/r2z_capture.log 2>&1
tail -30 gpurun_out/r2z_capture.log | cut -c1-200
