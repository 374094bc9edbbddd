#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "balanced" > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log; grep -E "^E  " gpurun_out/r2m_pytest.log | head -20
for B in "" "--balance-evals"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 --regions 3 --workload cfg3_sphere_256x512 --mode nearest --only-mode --denoise-us 500 --no-parity $B 2>gpurun_out/r2m_err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$B', round(d['ms_per_step']*1e3,1),'us', d['config']['cuda_graph'], d['config']['balance_evals'])"
tail -2 gpurun_out/r2m_err.log
done
