#!/bin/bash
# whole -m gpu suite, the reference's World-API unit tests on CudaWorld, the bench line
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2t_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2t_pytest_gpu.log
timeout 1500 python tools/run_unittests_cuda.py --timeout 60 > gpurun_out/r2t_unittests_cuda.txt 2> gpurun_out/r2t_unittests_cuda.err
head -34 gpurun_out/r2t_unittests_cuda.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2t.json 2> gpurun_out/r2t.err
python -c "
import json;d=json.load(open('gpurun_out/r2t.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'], d['roofline']['frac'])"
