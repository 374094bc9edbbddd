#!/bin/bash
# ncu evidence for the hot kernels (run through gpurun, 1 GPU).
#   launches_<tag>.csv : every launch with its device time (cold-cache, serialised)
#   prof_<tag>.ncu-rep : --set full capture of the gather / merge kernels (one bilinear + one nearest step)
TAG=${1:-r1}
LAYOUT=${2:-pm}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu --no-kernel-pass --layout $LAYOUT > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_|merge_' -s 12 -c 4 \
    -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 4 --warmup 3 --no-graph --no-cpu --no-kernel-pass --layout $LAYOUT > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
