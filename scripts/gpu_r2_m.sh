#!/bin/bash
# Round 2, call M (N GPUs): the multi-GPU path as a product feature and as the driver's scaling bench.
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== entry point on $N GPUs (NCCL gather of the final frames on rank 0)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/multigpu_entry_check.py 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -4
echo "=== bench, N = $N (reference arm, own arm)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>/dev/null | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_m.err | tee gpurun_out/r02_bench_${N}gpu_v1.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d.get('torch_cuda_baseline',{}); print('N', d['n_gpus'], 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],3), 'torch', t.get('value'), 'vs_torch', d.get('vs_torch_cuda'), 'finite', d['config']['finite'])"
tail -3 gpurun_out/bench_m.err
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N 1 (same box) value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
