#!/bin/bash
# Round 2, call B: whole GPU suite (new full-size parity tests), warm ncu launch lists (default / GN epilogue stats), bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== whole GPU suite"
timeout 1500 python -m pytest tests/ -q -m gpu --durations=12 -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/r02_pytest_gpu_all_v1.log
echo "=== launch list (warm, eager one step)"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v1_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
CGD_GN_EPI_STATS=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v1_epi_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches_epi.log 2>&1; tail -2 gpurun_out/ncu_launches_epi.log
echo "=== bench (with PyTorch-CUDA and CPU arms)"
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_v1.json | cut -c1-1500
tail -5 gpurun_out/bench.err
