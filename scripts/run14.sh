bash scripts/gpu_tests.sh tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py 2>&1 | grep -v "^$" | tail -40
timeout 300 python scripts/conv_microbench.py 2>&1 | head -5
CGD_CONV_HALO=0 timeout 300 python scripts/conv_microbench.py 2>&1 | head -5
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-250
