#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
S='[{},{"tile_w":16},{"tile_w":32},{"tile_w":8},{"merge_minb":5},{"tile_w":16,"merge_minb":5}]'
KBENCH_NO_DENOISE=1 KBENCH_SETTINGS="$S" timeout 200 python tools/kbench.py cfg2_sphere_128x256 2>&1 | tee gpurun_out/r2a_kbench_cfg2.jsonl | cut -c1-600
S='[{},{"tile_w":16},{"tile_w":32},{"tile_w":8}]'
KBENCH_NO_DENOISE=1 KBENCH_SETTINGS="$S" timeout 200 python tools/kbench.py cfg3_sphere_256x512 2>&1 | tee gpurun_out/r2a_kbench_cfg3.jsonl | cut -c1-600
DS_TUNE="4=16,5=5" timeout 200 python -m pytest tests/test_gpu_pm.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2a_pytest_tuned.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2a_pytest.log
