python scripts/gn_microbench.py 2>&1 | grep -E "grid|twopass" | tail -16
bash scripts/gpu_bench_only.sh > gpurun_out/bench_only.log 2>&1; cat gpurun_out/bench.json | cut -c1-200
