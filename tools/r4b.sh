#!/bin/bash
# round-2 late check: sharded planar ring loop on 2 GPUs
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_multi.py -q -x -k "sharded_i2v_ring_denoise_loop or sharded_ring_denoise_loop" 2>&1 | tail -30 | tee gpurun_out/r4b_pytest.log
