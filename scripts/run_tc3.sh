#!/bin/bash
bash scripts/gpu_tests.sh tests/test_gpu_conv.py tests/test_gpu_ops.py
echo "=== bench cluster split-K on"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_on.err | tee gpurun_out/bench_cluster_on.json | cut -c1-330; tail -3 gpurun_out/bench_on.err
echo "=== bench cluster split-K off"; CGD_CONV_CLUSTER=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_cluster_off.json | cut -c1-330
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches_warm.log 2>&1
tail -2 gpurun_out/ncu_launches_warm.log; wc -l gpurun_out/launches_warm.csv
