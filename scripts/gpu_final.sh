#!/bin/bash
# full validation: every GPU test file, smoke(), the default bench (+ cold/warm launch lists), then the extra workloads
bash scripts/gpu_tests.sh tests/test_gpu_*.py
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
bash scripts/gpu_bench_only.sh
for w in cfg3 cfg4 cfg5; do
  echo "=== workload $w"
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w 2> gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json
  tail -3 gpurun_out/bench_$w.err
done
