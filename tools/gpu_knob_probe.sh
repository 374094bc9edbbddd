#!/bin/bash
# A/B of a ds_tune knob on a B200 (run through gpurun): kernel timings for each setting in
# KBENCH_SETTINGS, then the pixel-major parity tests under DS_TUNE (results must not change).
#   usage: gpu_knob_probe.sh '<KBENCH_SETTINGS json>' '<DS_TUNE knob=value,...>' [tag]
mkdir -p gpurun_out
TAG=${3:-knob}
KBENCH_NO_DENOISE=1 KBENCH_SETTINGS="$1" timeout 120 python tools/kbench.py 2>&1 | tee gpurun_out/kbench_$TAG.jsonl | cut -c1-420
DS_TUNE="$2" timeout 80 python -m pytest tests/test_gpu_pm.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_$TAG.log
