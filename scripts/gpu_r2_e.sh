#!/bin/bash
# Round 2, call E: same-box A/B of the round-2 kernel changes (box-to-box variance is ~5 %: compare within one call only).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== tests: narrow conv, QuickGELU epilogue, special layouts, streaming GroupNorm engine forced, guidance"
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_norm.py tests/test_gpu_guidance.py -q -m gpu -x --tb=short -p no:cacheprovider -k "narrow or quick_gelu or special or (forced and stream) or guidance or cutout or augs" 2>&1 | tail -6
for e in direct stream; do
  echo "=== gn microbench, engine $e"
  CGD_GN_GRID_ENGINE=$e GN_ONLY=grid timeout 300 python scripts/gn_microbench.py 2>&1 | grep -E "HW +(4096|16384|65536)" | tee gpurun_out/r02_gn_microbench_v2_$e.txt
done
b() { echo "--- $1"; env $1 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],3), 'dom_us', round(d['roofline']['avg_launch_s']*1e6,1), 'conv_ms', round(d['roofline']['step_conv_ms'],2), 'launches', d['launches_per_step'])"; }
echo "=== same-box A/B (image-steps/s)"
b "CGD_NOP=1"
b "CGD_GN_GRID_ENGINE=stream"
b "CGD_GN_GRID_ENGINE=stream CGD_GN_EPI_STATS=1"
b "CGD_QGELU_EPI=0"
b "CGD_CONV_NARROW=0"
b "CGD_QGELU_EPI=0 CGD_CONV_NARROW=0"
b "CGD_NOP=2"
echo "=== launch list (default)"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v3_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
CGD_GN_GRID_ENGINE=stream timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v3_stream_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches2.log 2>&1; tail -1 gpurun_out/ncu_launches2.log
