#!/bin/bash
# First GPU call of the next round: what was built after this round's GPU budget ran out.
#   1. the ModifiedResNet tower on the device (opt-in tests), 2. the whole suite as the driver runs it, 3. the bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== RN tower (CGD_TEST_RN=1)"
CGD_TEST_RN=1 timeout 900 python -m pytest tests/test_gpu_rn.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/test_gpu_rn.log
echo "=== GroupNorm from conv-epilogue statistics (CGD_TEST_EPI=1), then the step with the path on"
CGD_TEST_EPI=1 timeout 600 python -m pytest tests/test_gpu_epi_stats.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -20 | tee gpurun_out/test_gpu_epi.log
CGD_GN_EPI_STATS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-330
echo "=== split last wave of the pair conv kernel (CGD_TEST_TAIL=1), dominant layer alone both ways, step with it on"
CGD_TEST_TAIL=1 timeout 600 python -m pytest tests/test_gpu_tail.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -20 | tee gpurun_out/test_gpu_tail.log
timeout 300 python scripts/conv_microbench.py 2>&1 | tail -12
CGD_CONV_TAIL=1 timeout 300 python scripts/conv_microbench.py 2>&1 | tail -12
CGD_CONV_TAIL=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-330
echo "=== whole GPU suite, one process"
timeout 1700 python -m pytest tests/ -x -q -m gpu --durations=8 -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_all.log
echo "=== bench"
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
