#!/bin/bash
# ncu --set full captures of the dominant kernels of one eager cfg2 step (one invocation per kernel family); raw pages as CSV
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
    python scripts/profile_step.py eager > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
}
cap conv3x3_256 conv_tc2_kernel 1 2
cap gn_grid 'gn_(fwd|bwd)_grid' 0 2
cap gn_grid_bwd 'gn_bwd_grid' 20 2
ls -la gpurun_out/*.ncu-rep
