#!/bin/bash
# Round 2, call N (last call of the round): the 128x128 checkpoint (wide-head attention, csrc/attention_wide.cu) and the RN50x4 / RN50x16
# towers (zero-padded widths) on the device -- new cases first, then the rest of the suite, smoke, the two bench lines.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
FIRST="tests/test_gpu_attention.py tests/test_gpu_norm.py tests/test_gpu_rn.py tests/test_gpu_baseline_configs.py tests/test_gpu_conv.py"
REST=$(ls tests/test_gpu_*.py | grep -v -e test_gpu_attention.py -e test_gpu_norm.py -e test_gpu_rn.py -e test_gpu_baseline_configs.py -e test_gpu_conv.py | tr '\n' ' ')
echo "=== pytest -m gpu (new cases first, no -x)"
timeout 620 python -m pytest $FIRST $REST -q -m gpu --durations=8 -p no:cacheprovider --timeout 240 -s -rf > gpurun_out/r02_pytest_gpu_all_v4_full.log 2>&1
grep -v "^$" gpurun_out/r02_pytest_gpu_all_v4_full.log | grep -e "cfg[2-5] \|default128 \|passed\|failed\|FAILED\|Error\|error" | tail -40 | tee gpurun_out/r02_pytest_gpu_all_v4.log
echo "=== smoke()"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench default128"
timeout 240 python bench.py --workload default128 --steps 20 --warmup 5 2> gpurun_out/bench128.err | tee gpurun_out/r02_bench_default128_v1.json | cut -c1-400
echo "=== bench cfg2"
timeout 240 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_v5.json | cut -c1-300
tail -3 gpurun_out/bench128.err
