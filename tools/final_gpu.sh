#!/bin/bash
# round-end GPU pass: parity tests, the three bench lines (mixed 1M, tumbler arena, reference arm)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python bench.py --workload tumblers --steps 20 --warmup 3 > gpurun_out/bench_tumblers_final.json 2> gpurun_out/bench_tumblers_final.err
python -c "
import json;d=json.load(open('gpurun_out/bench_tumblers_final.json'));print(d['ms_per_step'],d['config']['colors'],d.get('cpu_baseline'))"
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_mixed_final2.json 2> gpurun_out/bench_mixed_final2.err
python -c "
import json;d=json.load(open('gpurun_out/bench_mixed_final2.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['clocks'])"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
cut -c1-600 gpurun_out/bench_ref_final.json
