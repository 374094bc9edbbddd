#!/bin/bash
# A/B of the exchange ordering protocol (host sync kernels vs deferred in-kernel publication) at N GPUs,
# with the parity tests under both.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "two_gpus or transports or denoise_loop" > gpurun_out/r2g_pytest_host_n$N.log 2>&1; tail -3 gpurun_out/r2g_pytest_host_n$N.log
DS_TEST_KERNEL_PUBLISH=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "two_gpus or transports or denoise_loop" > gpurun_out/r2g_pytest_deferred_n$N.log 2>&1; tail -3 gpurun_out/r2g_pytest_deferred_n$N.log
for PUB in kernel host; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --merge-publish $PUB 2>gpurun_out/r2g_bench_${PUB}_n${N}_err.log | tee gpurun_out/r2g_bench_${PUB}_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$PUB', round(d['ms_per_step']*1e3,2),'us', {k:round(v['ms_per_step']*1e3,2) for k,v in d['variants'].items()}, d['parity']['ok'], 'e2e', round(d['e2e']['ms_per_step'],3))"
  tail -2 gpurun_out/r2g_bench_${PUB}_n${N}_err.log
done
