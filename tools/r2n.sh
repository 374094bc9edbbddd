#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "balanced and 8" > gpurun_out/r2n_pytest.log 2>&1; tail -2 gpurun_out/r2n_pytest.log
for T in 2000 200; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 20 --warmup 3 --regions 5 --workload cfg3_sphere_256x512 --mode nearest --only-mode --denoise-us $T --no-parity --balance-evals 2>gpurun_out/r2n_t${T}_err.log > gpurun_out/r2n_cfg3_denoise${T}_balanced_n8.json
python -c "
import json
d=json.loads(open('gpurun_out/r2n_cfg3_denoise${T}_balanced_n8.json').read().strip().splitlines()[-1]); print('T=$T balanced N=8:', round(d['ms_per_step']*1e3,1),'us/step graph', d['config']['cuda_graph'], d['config']['balance_evals'][:40])"
tail -2 gpurun_out/r2n_t${T}_err.log | cut -c1-200
done
