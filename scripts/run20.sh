bash scripts/gpu_tests.sh tests/test_gpu_norm.py tests/test_gpu_attention.py tests/test_gpu_ops.py 2>&1 | grep -v "^$" | tail -24
python scripts/gn_microbench.py 2>&1 | grep -E "grid" | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-250
CGD_CONV_FUSE_REDUCE=1 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-200
