#!/bin/bash
# 1-GPU round check: smoke, GPU parity tests (incl. the full-size digests), default bench + reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/r2b_bench_err.log | tee gpurun_out/r2b_bench.json | cut -c1-1500
tail -3 gpurun_out/r2b_bench_err.log
timeout 300 python bench.py --steps 20 --warmup 3 --workload cfg3_sphere_256x512 --no-cpu 2>gpurun_out/r2b_bench_cfg3_err.log | tee gpurun_out/r2b_bench_cfg3.json | cut -c1-1200
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tee gpurun_out/r2b_bench_ref.json | cut -c1-800
