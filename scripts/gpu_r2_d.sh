#!/bin/bash
# Round 2, call D: streaming GroupNorm engine (norm_stream.cu) + row-based cutout kernels: correctness, microbench per engine, step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== norm / epi / guidance / ops tests"
timeout 1200 python -m pytest tests/test_gpu_norm.py tests/test_gpu_epi_stats.py tests/test_gpu_guidance.py tests/test_gpu_ops.py tests/test_gpu_baseline_configs.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -25
for e in ring direct stream; do
  echo "=== gn microbench, engine $e"
  CGD_GN_GRID_ENGINE=$e GN_ONLY=grid timeout 300 python scripts/gn_microbench.py 2>&1 | grep -E "HW +(4096|16384|65536)" | tee gpurun_out/r02_gn_microbench_$e.txt
done
echo "=== bench default"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | tee gpurun_out/r02_bench_v2_stream.json | cut -c1-330
echo "=== bench EPI"
CGD_GN_EPI_STATS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | tee gpurun_out/r02_bench_v2_stream_epi.json | cut -c1-330
echo "=== bench persistent engines (A/B)"
CGD_GN_GRID_ENGINE=direct timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | cut -c1-330
echo "=== co-resident pair kernel: conv tests, then the step with it on"
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x --tb=short -p no:cacheprovider -k co_resident 2>&1 | tail -5
CGD_CONV_CO=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | tee gpurun_out/r02_bench_v2_stream_co.json | cut -c1-330
CGD_CONV_CO=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v2_stream_co_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches3.log 2>&1; tail -1 gpurun_out/ncu_launches3.log
echo "=== launch list (stream + EPI)"
CGD_GN_EPI_STATS=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v2_stream_epi_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v2_stream_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches2.log 2>&1; tail -2 gpurun_out/ncu_launches2.log
