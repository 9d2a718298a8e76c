#!/bin/bash
# Round 2, call F: one-MUFU sigmoid (tanh.approx) in every SiLU / SiLU': accuracy on the parity tests, same-box A/B against the ex2 + rcp build.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== parity with the tanh sigmoid: norm, ops, full-size configs, chains"
timeout 1500 python -m pytest tests/test_gpu_norm.py tests/test_gpu_ops.py tests/test_gpu_baseline_configs.py tests/test_gpu_chain.py tests/test_gpu_fullsize.py tests/test_gpu_guidance.py tests/test_gpu_lpips.py -q -m gpu -x --tb=short -p no:cacheprovider -s 2>&1 | grep -E "^cfg|passed|failed|Error|error" | tail -12
for e in direct stream; do
  echo "=== gn microbench (tanh sigmoid), engine $e"
  CGD_GN_GRID_ENGINE=$e GN_ONLY=grid timeout 300 python scripts/gn_microbench.py 2>&1 | grep -E "HW +(4096|16384|65536)" | tee gpurun_out/r02_gn_microbench_v3_$e.txt
done
b() { echo "--- $1"; env $1 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],3), 'dom_us', round(d['roofline']['avg_launch_s']*1e6,1), 'conv_ms', round(d['roofline']['step_conv_ms'],2), 'launches', d['launches_per_step'])"; }
echo "=== same-box A/B, tanh build"
b "CGD_NOP=1"
b "CGD_GN_GRID_ENGINE=stream"
b "CGD_GN_GRID_ENGINE=stream CGD_GN_EPI_STATS=1"
b "CGD_GN_EPI_STATS=1"
echo "=== launch lists, tanh build"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v4_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
CGD_GN_GRID_ENGINE=stream CGD_GN_EPI_STATS=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v4_stream_epi_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches2.log 2>&1; tail -1 gpurun_out/ncu_launches2.log
echo "=== rebuild with the ex2 + rcp sigmoid (-DCGD_SIGMOID_EX2RCP), same box"
CGD_NVCC_EXTRA=-DCGD_SIGMOID_EX2RCP python -c "from clip_guided_diffusion_b200 import build as b; b.build(force=True)" > gpurun_out/build2.log 2>&1 || tail -5 gpurun_out/build2.log
b "CGD_NOP=2"
b "CGD_GN_GRID_ENGINE=stream CGD_GN_EPI_STATS=1"
