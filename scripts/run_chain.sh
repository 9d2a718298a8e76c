#!/bin/bash
bash scripts/gpu_tests.sh tests/test_gpu_norm.py tests/test_gpu_chain.py tests/test_gpu_ops.py tests/test_gpu_guidance.py tests/test_gpu_fullsize.py
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_nocpu.json | cut -c1-330
