#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2z_pytest.log
KBENCH_SETTINGS='[{},{"hw_const":0}]' timeout 300 python tools/kbench.py cfg2_sphere_128x256 2>&1 | tee gpurun_out/r2z_kbench_cfg2.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:v for k,v in d.items() if k.endswith('_us') or k=='hw_const'})"
bash tools/gpu_profile.sh r2c > gpurun_out/r2z_profile.log 2>&1; tail -2 gpurun_out/r2z_profile.log
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/r2z_bench_err.log > gpurun_out/r2z_bench.json; tail -2 gpurun_out/r2z_bench_err.log
python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print(round(d['ms_per_step']*1e3,2), r['kernel'], round(r['frac'],3), round(r['frac_algorithmic'],3), round(r['step_frac'],3), round(r['step_frac_algorithmic'],3), r['plain_copy_same_bytes'], d['variants'])"
