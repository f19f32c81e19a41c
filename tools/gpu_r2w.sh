#!/bin/bash
# whole -m gpu suite, the World-API unit tests on CudaWorld, default bench run (as the driver runs it: extras + cpu baseline)
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2w_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2w_pytest_gpu.log
timeout 1500 python tools/run_unittests_cuda.py --timeout 60 > gpurun_out/r2w_unittests_cuda.txt 2> gpurun_out/r2w_unittests_cuda.err
head -30 gpurun_out/r2w_unittests_cuda.txt | cut -c1-200
( time timeout 900 python bench.py ) > gpurun_out/r2w_bench_default.json 2> gpurun_out/r2w_bench_default.err
tail -4 gpurun_out/r2w_bench_default.err
python -c "
import json;d=json.loads(open('gpurun_out/r2w_bench_default.json').read().strip().splitlines()[0]);print(d['ms_per_step'],d['e2e']['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['roofline']['frac'], d.get('tumblers_n1',{}).get('ms_per_step'), d.get('toi_serial_exact'), d['cpu_baseline'])"
( time timeout 600 python bench.py --impl reference ) 2>&1 | tail -5 | cut -c1-300
