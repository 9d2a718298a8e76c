bash scripts/gpu_tests.sh tests/test_gpu_conv.py tests/test_gpu_norm.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py 2>&1 | grep -v "^$" | tail -24
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-250
CGD_CONV_FUSE_REDUCE=0 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-200
