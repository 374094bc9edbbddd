#!/bin/bash
# Gather tile-order experiment (DS_TUNE_PM_GATHER_BLOCKED): kernel timings, then parity under the knob.
mkdir -p gpurun_out
KBENCH_NO_DENOISE=1 KBENCH_SETTINGS='[{"blocked":0},{"blocked":3},{"blocked":3,"gather_cps":3},{"blocked":3,"gather_cps":2},{"blocked":0,"gather_cps":3}]' \
  timeout 120 python tools/kbench.py 2>&1 | tee gpurun_out/kbench_blocked.jsonl | cut -c1-400
DS_TUNE="3=3" timeout 60 python -m pytest tests/test_gpu_pm.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_blocked.log
