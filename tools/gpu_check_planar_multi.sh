#!/bin/bash
# Sharded planar loops on a 2-GPU box (through gpurun --gpus 2).
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_multi.py -q -x -k "sharded_i2v_ring_denoise_loop or sharded_ring_denoise_loop" 2>&1 | tail -30 | tee gpurun_out/planar_multi_pytest.log
