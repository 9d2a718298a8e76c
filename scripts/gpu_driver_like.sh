#!/bin/bash
# what the driver runs at round end: the whole GPU suite in ONE process, then smoke()
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1700 python -m pytest tests/ -x -q -m gpu --durations=8 -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
