#!/bin/bash
# Round 2, call Q: ViT-L/14@336px (336 px cutouts, 577 tokens) full guided step vs the fp32 oracle; final-state lines of two other workloads.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 200 python -m pytest tests/test_gpu_baseline_configs.py -q -m gpu -p no:cacheprovider -s -k "vit_l14_336 or cfg2" 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/r02_pytest_vit336_v1.log
for w in cfg3 cfg4; do
  timeout 100 python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | cut -c1-1200 >> gpurun_out/r02_bench_workloads_v3.txt
done
cut -c1-260 gpurun_out/r02_bench_workloads_v3.txt
