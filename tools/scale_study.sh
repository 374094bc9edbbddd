#!/bin/bash
# Multi-GPU evidence on an N-GPU lease (run through `gpurun --gpus N`):
#   1. the multi-GPU parity tests that fit the lease (-k filter in $PYTEST_K),
#   2. bench.py at every world size in $WORLDS on cfg2 (the headline) -- every line carries parity.ok,
#   3. cfg3 (north-star shape 256x512) at the largest world,
#   4. --denoise-us sweep at the largest world and at 1 GPU: a calibrated busy kernel per local window between
#      gather and merge, the stand-in for the UNet that cannot travel to the box (SCALE with the UNet in the loop).
WORLDS=${WORLDS:-"2 4 8"}
TAG=${TAG:-r2k}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv | sort | uniq -c
if [ -n "$PYTEST_K" ]; then
  timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "$PYTEST_K" > gpurun_out/${TAG}_pytest_multi.log 2>&1
  grep -E "passed|failed" gpurun_out/${TAG}_pytest_multi.log | tail -3
fi
run() {  # world tag args...
  local N=$1 NAME=$2; shift 2
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 "$@" 2>gpurun_out/${TAG}_${NAME}_n${N}_err.log > gpurun_out/${TAG}_${NAME}_n$N.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 "$@" 2>gpurun_out/${TAG}_${NAME}_n${N}_err.log > gpurun_out/${TAG}_${NAME}_n$N.json
  fi
  python - "$N" "$NAME" gpurun_out/${TAG}_${NAME}_n$N.json <<'PY'
import json, sys
n, name, f = sys.argv[1:]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f"N={n} {name}: {d['ms_per_step']*1e3:.2f} us/step, {d['value']:.0f} blends/s, e2e {d['e2e']['ms_per_step']:.3f} ms, parity {d.get('parity', {}).get('ok')}")
except Exception as e:
    print(f"N={n} {name}: FAILED ({e})")
PY
}
LAST=$(echo $WORLDS | awk '{print $NF}')
for N in $WORLDS; do run $N cfg2; done
run $LAST cfg3 --workload cfg3_sphere_256x512 --no-cpu
for T in 200 2000; do
  run $LAST cfg3_denoise$T --workload cfg3_sphere_256x512 --mode nearest --only-mode --denoise-us $T --no-parity --regions 5
  run 1 cfg3_denoise$T --workload cfg3_sphere_256x512 --mode nearest --only-mode --denoise-us $T --no-cpu --regions 5
done
