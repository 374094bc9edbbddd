#!/bin/bash
# N-GPU check: multi-GPU parity tests (worlds that fit the lease), then the bench at N GPUs with parity.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv | head -3
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/r2e_pytest_multi_n$N.log 2>&1
grep -E "Error|error|assert|Traceback|File \"/root" gpurun_out/r2e_pytest_multi_n$N.log | head -30
tail -6 gpurun_out/r2e_pytest_multi_n$N.log
[ -n "$SKIP_BENCH" ] && exit 0
for WL in cfg2_sphere_128x256 cfg1_ring_64x128; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --workload $WL 2>gpurun_out/r2e_bench_${WL}_n${N}_err.log | tee gpurun_out/r2e_bench_${WL}_n$N.json | cut -c1-1800
  tail -3 gpurun_out/r2e_bench_${WL}_n${N}_err.log
done
