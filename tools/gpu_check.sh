#!/bin/bash
# Round check on a B200 box (run through gpurun): smoke, GPU parity tests, bench (both layouts).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
for layout in ${LAYOUTS:-pm ref}; do
  python bench.py --steps 200 --warmup 20 --layout $layout 2>gpurun_out/bench_err_$layout.log | tee gpurun_out/bench_$layout.json
  tail -5 gpurun_out/bench_err_$layout.log
done
