#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2i_pytest.log
bash tools/gpu_profile.sh r2a > gpurun_out/r2i_profile.log 2>&1; tail -3 gpurun_out/r2i_profile.log
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/r2i_bench_err.log | tee gpurun_out/r2i_bench.json | cut -c1-600
