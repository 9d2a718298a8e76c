#!/bin/bash
# Round 2, call W: in-cluster split-K also for one-wave long-K layers (CGD_CLUSTER_WIDE=1): parity with it on, same-box A/B, and the
# final full suite (the persistent conv grid now comes from the device's SM count).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== parity with CGD_CLUSTER_WIDE=1 (cfg2 / cfg4 full-size step vs oracle, UNet properties)"
CGD_CLUSTER_WIDE=1 timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_fullsize.py -q -m gpu -x --tb=short -p no:cacheprovider -s -k "cfg2 or cfg4 or cfg1" 2>&1 | grep -E "^cfg|passed|failed|rror" | tail -6
b() { echo "--- $1"; env $1 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],3), 'dom_us', round(d['roofline']['avg_launch_s']*1e6,1), 'conv_ms', round(d['roofline']['step_conv_ms'],2), 'launches', d['launches_per_step'])"; }
echo "=== same-box A/B"
b "CGD_NOP=1"
b "CGD_CLUSTER_WIDE=1"
b "CGD_NOP=2"
b "CGD_CLUSTER_WIDE=1"
CGD_CLUSTER_WIDE=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v6_wide_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
echo "=== whole GPU suite (default settings)"
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
