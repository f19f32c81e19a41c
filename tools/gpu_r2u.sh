#!/bin/bash
# the driver's scaling launch at N=2: sharded tumbler worlds (strong scaling), NCCL for the statistics
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2u_n2.json 2> gpurun_out/r2u_n2.err
tail -3 gpurun_out/r2u_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/r2u_n2.json').read().strip().splitlines()[-1]);print(d['n_gpus'], d['ms_per_step'], round(d['value']/1e6,1), d['scaling'], d['config']['workload'][:80], d.get('gpu_launches'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2u_ref2.json 2> gpurun_out/r2u_ref2.err
tail -2 gpurun_out/r2u_ref2.json | cut -c1-400
