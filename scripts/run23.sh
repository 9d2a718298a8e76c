bash scripts/gpu_tests.sh tests/test_gpu_ops.py tests/test_gpu_guidance.py 2>&1 | grep -v "^$" | tail -30
