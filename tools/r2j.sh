#!/bin/bash
mkdir -p gpurun_out
KBENCH_NO_DENOISE=1 KBENCH_DEFAULT_ONLY=1 timeout 200 python tools/kbench.py cfg2_sphere_128x256 2>&1 | tee gpurun_out/r2j_kbench_cfg2.jsonl | cut -c1-500
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2j_pytest.log
bash tools/gpu_profile.sh r2b > gpurun_out/r2j_profile.log 2>&1; tail -2 gpurun_out/r2j_profile.log
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/r2j_bench_err.log | tee gpurun_out/r2j_bench.json | cut -c1-400
