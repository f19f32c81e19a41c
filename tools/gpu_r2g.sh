#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -4 gpurun_out/r2g_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2g_bench.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac']);print(d.get('toi_serial_exact'));print(d.get('tumblers_n1'));print(d.get('cpu_baseline'));print({k:(round(v['ms'],3),round(v['frac'],3)) for k,v in d['roofline']['other_kernels'].items()})"
