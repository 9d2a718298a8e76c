#!/bin/bash
# Run the GPU test files in separate processes (a trapped kernel poisons only its own process), logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in "$@"; do
  name=$(basename $t .py)
  echo "=== $t (CGD_TEST_CONV_IMPL=${CGD_TEST_CONV_IMPL:-0})"
  timeout 1200 python -m pytest $t -q -m gpu --tb=short --no-header -p no:cacheprovider 2>&1 | tail -n 120 | tee gpurun_out/${name}_impl${CGD_TEST_CONV_IMPL:-0}.log
done
