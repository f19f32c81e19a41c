#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus ${NG:-2} --steps 10 --warmup 3 > gpurun_out/r3i_bench_${NG:-2}gpu.json 2> gpurun_out/r3i_bench_${NG:-2}gpu.err
tail -2 gpurun_out/r3i_bench_${NG:-2}gpu.err | cut -c1-200
python -c "
import json;d=json.loads(open('gpurun_out/r3i_bench_${NG:-2}gpu.json').read().strip().splitlines()[-1]);print(d['n_gpus'], d['ms_per_step'], d['value'], d['scaling'], d['config']['workload'][:80])"
