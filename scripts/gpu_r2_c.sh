#!/bin/bash
# Round 2, call C: evidence.  ncu --set full over one launch of every kernel family; compute-sanitizer over small kernel tests.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== ncu --set full, one launch per kernel family"
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_families python scripts/profile_families.py > gpurun_out/ncu_families.log 2>&1
tail -4 gpurun_out/ncu_families.log
ncu -i gpurun_out/r02_families.ncu-rep --page raw --csv > gpurun_out/r02_families_raw.csv 2>/dev/null; wc -l gpurun_out/r02_families_raw.csv
echo "=== ncu --set full: conv with epilogue statistics + one-trip GroupNorm apply (CGD_GN_EPI_STATS=1)"
CGD_GN_EPI_STATS=1 PF_ONLY=GN_APPLY_EPI,CONV_STATS timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_epi python scripts/profile_families.py > gpurun_out/ncu_epi.log 2>&1
tail -3 gpurun_out/ncu_epi.log
ncu -i gpurun_out/r02_epi.ncu-rep --page raw --csv > gpurun_out/r02_epi_raw.csv 2>/dev/null
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --log-file gpurun_out/r02_sanitizer_$tool.log \
    python -m pytest tests/test_gpu_conv.py tests/test_gpu_norm.py tests/test_gpu_attention.py -q -m gpu -x -p no:cacheprovider \
      -k "(conv3x3_64x64_c128 or conv3x3_32x32_c256_res or conv3x3_16x16_c512_splitk or conv1x1_skip or linear_m50 or cluster8 or special or norm or attention) and not simt and not tc1" 2>&1 | tail -4
  echo "exit $?"; tail -5 gpurun_out/r02_sanitizer_$tool.log
done
