#!/bin/bash
# dataflow solver (k_flow_pass): parity, then A/B against the per-colour launches (debug bit 7)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_events.py tests/test_batch.py tests/test_set_contact.py -m gpu -x -q ) 2>&1 | tail -5
for f in 0 128; do
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --debug-flags $f > gpurun_out/r2m_$f.json 2> gpurun_out/r2m_$f.err
python -c "
import json;d=json.load(open('gpurun_out/r2m_$f.json'));print($f, d['ms_per_step'],{k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['gpu_launches'])"
done
