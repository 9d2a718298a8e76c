bash scripts/gpu_tests.sh tests/test_gpu_norm.py tests/test_gpu_guidance.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py 2>&1 | grep -v "^$" | tail -24
python scripts/gn_microbench.py 2>&1 | grep -E "grid" | tail -12
bash scripts/gpu_bench_only.sh > gpurun_out/bench_only.log 2>&1; cat gpurun_out/bench.json | cut -c1-300
