#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r3b_pytest.log
bash tools/gpu_profile.sh r2h > gpurun_out/r3b_profile.log 2>&1; tail -2 gpurun_out/r3b_profile.log
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/r3b_bench_err.log > gpurun_out/r3b_bench.json; tail -2 gpurun_out/r3b_bench_err.log
timeout 300 python bench.py --steps 20 --warmup 3 --workload cfg3_sphere_256x512 --no-cpu 2>>gpurun_out/r3b_bench_err.log > gpurun_out/r3b_bench_cfg3.json
python -c "
import json
for f in ('gpurun_out/r3b_bench.json','gpurun_out/r3b_bench_cfg3.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(d['config']['workload'], round(d['ms_per_step']*1e3,2), r['kernel'], 'frac', round(r['frac'],3), 'alg', round(r['frac_algorithmic'],3), 'step', round(r['step_frac'],3), round(r['step_frac_algorithmic'],3), {m:{k:(round(x['us'],2), round(x.get('kernel_vs_plain_copy',0),2)) for k,x in v.items()} for m,v in r['all'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3))"
