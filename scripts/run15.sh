bash scripts/gpu_tests.sh tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py 2>&1 | grep -v "^$" | tail -20
for s in 1 2 4; do echo "== CGD_VIT_STREAMS=$s"; CGD_VIT_STREAMS=$s timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-200; done
