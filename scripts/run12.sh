bash scripts/gpu_tests.sh tests/test_gpu_norm.py tests/test_gpu_attention.py tests/test_gpu_conv.py tests/test_gpu_ops.py tests/test_gpu_guidance.py tests/test_gpu_fullsize.py 2>&1 | grep -v "^$" | tail -40
python scripts/gn_microbench.py 2>&1 | grep -E "grid|fused" | tail -24
bash scripts/gpu_bench_only.sh > gpurun_out/bench_only.log 2>&1; cat gpurun_out/bench.json | cut -c1-300
CGD_CONV_PREFETCH=0 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-200
