#!/bin/bash
# Planar-loop check on a 1-GPU box (through gpurun): the three planar reference loops, the ring regression subset, smoke.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ring_loop.py -q 2>&1 | tail -30 | tee gpurun_out/planar_pytest.log
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full.py tests/test_gpu_vae_tiling.py -q -k "ring or cfg1 or cfg5 or wrap" 2>&1 | tail -8 | tee gpurun_out/planar_pytest_ring.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/planar_smoke.log
