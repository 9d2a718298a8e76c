#!/bin/bash
# Round 2, call A: device-validate what round 1 wrote but never ran (RN tower, conv-epilogue GN statistics, split last wave).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "=== RN tower (CGD_TEST_RN=1)"
CGD_TEST_RN=1 timeout 600 python -m pytest tests/test_gpu_rn.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/test_gpu_rn.log
echo "=== GN epilogue stats (CGD_TEST_EPI=1)"
CGD_TEST_EPI=1 timeout 400 python -m pytest tests/test_gpu_epi_stats.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -20 | tee gpurun_out/test_gpu_epi.log
echo "=== tail (CGD_TEST_TAIL=1)"
CGD_TEST_TAIL=1 timeout 400 python -m pytest tests/test_gpu_tail.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -20 | tee gpurun_out/test_gpu_tail.log
echo "=== unsynced loop (staging ring)"
timeout 400 python -m pytest tests/test_gpu_chain.py -q -m gpu -x --tb=short -p no:cacheprovider -k sync 2>&1 | tail -8
echo "=== microbench default / tail"
timeout 200 python scripts/conv_microbench.py 2>&1 | tail -12
CGD_CONV_TAIL=1 timeout 200 python scripts/conv_microbench.py 2>&1 | tail -12
echo "=== bench default"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-400
echo "=== bench EPI"
CGD_GN_EPI_STATS=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-400
echo "=== bench TAIL"
CGD_CONV_TAIL=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-400
echo "=== bench both"
CGD_GN_EPI_STATS=1 CGD_CONV_TAIL=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-400
