#!/bin/bash
# ncu evidence for the planar ring kernels at 4K scale (BASELINE config 5), 1 GPU.
TAG=${1:-ring}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --workload cfg5_ring_512x1024 --steps 6 --warmup 3 --no-graph --no-cpu --no-kernel-pass > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ring_copy|merge_nearest' -s 6 -c 2 \
    -f -o gpurun_out/prof_${TAG} \
    python bench.py --workload cfg5_ring_512x1024 --steps 3 --warmup 3 --no-graph --no-cpu --no-kernel-pass > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -4
