#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_norm.py
echo "=== microbench v2 (direct loads)"; GN_ONLY=grid timeout 600 python scripts/gn_microbench.py 2>&1 | grep -E " grid " | tee gpurun_out/gn_mb_v2.txt
echo "=== microbench ring"; CGD_GN_GRID_RING=1 GN_ONLY=grid timeout 600 python scripts/gn_microbench.py 2>&1 | grep -E " grid " | tee gpurun_out/gn_mb_ring.txt
echo "=== bench v2"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-330
echo "=== bench ring"; CGD_GN_GRID_RING=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-330
