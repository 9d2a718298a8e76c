#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=index,name,memory.used --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_m.err | tee gpurun_out/r02_bench_${N}gpu_v1.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d.get('torch_cuda_baseline',{}); print('N', d['n_gpus'], 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],3), 'torch', t.get('value', t), 'vs_torch', d.get('vs_torch_cuda'), 'finite', d['config']['finite'])"
tail -2 gpurun_out/bench_m.err
