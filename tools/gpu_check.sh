#!/bin/bash
# Round check on a 1-GPU B200 box (run through gpurun): smoke, GPU parity tests, the default bench + reference arm,
# and one short bench line per other workload.
mkdir -p gpurun_out
TAG=${1:-check}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/${TAG}_bench_err.log | tee gpurun_out/${TAG}_bench.json | cut -c1-300
tail -3 gpurun_out/${TAG}_bench_err.log
for WL in cfg3_sphere_256x512 cfg4_sphere_f64 cfg1_ring_64x128 cfg5_ring_512x1024; do
  timeout 400 python bench.py --steps 10 --warmup 3 --regions 5 --workload $WL 2>gpurun_out/${TAG}_bench_${WL}_err.log > gpurun_out/${TAG}_bench_$WL.json
  python - $WL gpurun_out/${TAG}_bench_$WL.json <<'PY'
import json, sys
wl, f = sys.argv[1:]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(f"{wl}: {d['ms_per_step']*1e3:.1f} us/step {d['value']:.0f} blends/s; roofline {r.get('kernel')} frac {r.get('frac', 0):.2f} alg {r.get('frac_algorithmic', 0):.2f} step {r.get('step_frac', 0):.2f}/{r.get('step_frac_algorithmic', 0):.2f}; cpu {d.get('cpu_baseline', {}).get('value')}")
except Exception as e:
    print(wl, "FAILED", e)
PY
  tail -2 gpurun_out/${TAG}_bench_${WL}_err.log
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-300
