#!/bin/bash
# Round 2, call Z: the final state exactly as the driver runs it (suite, smoke, reference arm, own arm) after the last default flips.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== pytest -m gpu"
timeout 1700 python -m pytest tests/ -x -q -m gpu --durations=5 -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -18 | tee gpurun_out/r02_pytest_gpu_all_v3.log
echo "=== smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tee gpurun_out/r02_bench_reference_v3.json | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_v4.json | cut -c1-300
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v7_final_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
