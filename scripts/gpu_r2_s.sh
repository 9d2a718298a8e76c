#!/bin/bash
# Round 2, call S: the 8 x 8 weight-streaming conv kernel (conv_small_kernel): correctness, then same-box A/B in the step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== conv tests (8x8 cases, all impls), UNet-level parity"
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x --tb=short -p no:cacheprovider -k "8x8 or narrow or quick_gelu" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_fullsize.py tests/test_gpu_ops.py -q -m gpu -x --tb=short -p no:cacheprovider -s 2>&1 | grep -E "^cfg|passed|failed|rror" | tail -8
b() { echo "--- $1"; env $1 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],3), 'dom_us', round(d['roofline']['avg_launch_s']*1e6,1), 'conv_ms', round(d['roofline']['step_conv_ms'],2), 'launches', d['launches_per_step'])"; }
echo "=== same-box A/B"
b "CGD_NOP=1"
b "CGD_CONV_SMALL=0"
b "CGD_NOP=2"
echo "=== launch list"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v5s_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
