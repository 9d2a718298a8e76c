#!/bin/bash
# Round 2, call P: evidence for the 128x128 checkpoint path -- launch list of one eager step of the reference's default call, and
# compute-sanitizer (memcheck, racecheck) over the wide-head attention cases.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
CGD_PROFILE_WORKLOAD=default128 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_default128_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches128.log 2>&1; tail -1 gpurun_out/ncu_launches128.log
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 170 compute-sanitizer --tool $tool --error-exitcode 3 --log-file gpurun_out/r02_sanitizer_wide_$tool.log \
    python -m pytest tests/test_gpu_attention.py -q -m gpu -p no:cacheprovider -k "d128 or d192 or d256" 2>&1 | tail -2
  tail -3 gpurun_out/r02_sanitizer_wide_$tool.log
done
