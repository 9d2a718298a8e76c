bash scripts/gpu_tests.sh tests/test_gpu_lpips.py tests/test_gpu_norm.py tests/test_gpu_ops.py tests/test_gpu_guidance.py 2>&1 | grep -v "^$" | tail -40
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-250
