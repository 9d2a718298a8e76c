bash scripts/gpu_tests.sh tests/test_gpu_norm.py 2>&1 | grep -v "^$" | tail -30
python scripts/gn_microbench.py 2>&1 | grep -E "grid" | tail -12
timeout 900 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-260
