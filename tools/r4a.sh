#!/bin/bash
# round-2 late check: the planar ring loop tests + fused ring merge timing (1 GPU)
mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_gpu_ring_loop.py -q 2>&1 | tail -40 | tee gpurun_out/r4a_pytest.log
timeout 100 python tools/ring_ddim_bench.py --iters 20 2>gpurun_out/r4a_bench_err.log | tee gpurun_out/r4a_ring_ddim.jsonl
tail -3 gpurun_out/r4a_bench_err.log
