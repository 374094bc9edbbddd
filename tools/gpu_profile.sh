#!/bin/bash
# ncu evidence for the hot kernels (run through gpurun, 1 GPU).
#   launches_<tag>.csv      : every launch of a 20-step eager run with its device time (cold-cache, serialised)
#   prof_<tag>_<mode>.ncu-rep: --set full capture of one gather + one merge launch per mode
TAG=${1:-r2}
WL=${2:-cfg2_sphere_128x256}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu --no-kernel-pass --regions 1 --workload $WL > gpurun_out/ncu_bench_${TAG}.log 2>&1
# one gather + one merge launch per mode: skip the 3 warm-up steps (and, for the nearest mode, the 8 pack launches of
# make_state, which are pm_merge_nearest_kernel with the identity plan)
capture() {  # mode regex skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 2 \
      -f -o gpurun_out/prof_${TAG}_$1 \
      python bench.py --steps 4 --warmup 3 --no-graph --no-cpu --no-kernel-pass --regions 1 --mode $1 --only-mode --workload $WL > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
}
capture bilinear 'pm_gather_kernel|pm_merge_weighted' 6
capture nearest 'pm_gather_kernel|pm_merge_nearest' 14
ls -la gpurun_out/*${TAG}*
