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
echo "=== ncu --set full, one launch per kernel family (raw page only: the .ncu-rep files exceed the 64 MiB return limit)"
timeout 1200 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_families python scripts/profile_families.py > gpurun_out/ncu_families.log 2>&1
tail -3 gpurun_out/ncu_families.log
ncu -i /tmp/r02_families.ncu-rep --page raw --csv > gpurun_out/r02_families_raw.csv 2>/dev/null; ls -la gpurun_out/r02_families_raw.csv
CGD_GN_EPI_STATS=1 PF_ONLY=GN_APPLY_EPI,CONV_STATS timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_epi python scripts/profile_families.py > gpurun_out/ncu_epi.log 2>&1
ncu -i /tmp/r02_epi.ncu-rep --page raw --csv > gpurun_out/r02_epi_raw.csv 2>/dev/null
for tool in memcheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 --log-file gpurun_out/r02_sanitizer_$tool.log \
    python -m pytest tests/test_gpu_conv.py tests/test_gpu_norm.py tests/test_gpu_attention.py -q -m gpu -x -p no:cacheprovider \
      -k "(conv3x3_64x64_c128 or conv3x3_32x32_c256_res or conv3x3_16x16_c512_splitk or conv1x1_skip or linear_m50 or cluster8 or special or norm or attention) and not simt and not tc1 and not forced and not co_resident" 2>&1 | tail -3
  tail -3 gpurun_out/r02_sanitizer_$tool.log
done
du -sh gpurun_out
