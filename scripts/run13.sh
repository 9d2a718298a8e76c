bash scripts/gpu_tests.sh tests/test_gpu_norm.py tests/test_gpu_ops.py 2>&1 | grep -v "^$" | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --torch-baseline 2>/dev/null > gpurun_out/bench_torch.json; cut -c1-250 gpurun_out/bench_torch.json; python -c "import json; d=json.load(open('gpurun_out/bench_torch.json')); print(d.get('torch_cuda_baseline'))"
bash scripts/gpu_ncu_full.sh
