#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== bench $@"
timeout 1200 python bench.py --steps 20 --warmup 5 "$@" 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "=== ncu launch list (one step, eager launches; cold = recipe default, warm = --cache-control none)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches_warm.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv gpurun_out/launches_warm.csv
