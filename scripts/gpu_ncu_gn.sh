#!/bin/bash
# ncu --set full of the GroupNorm grid kernels at the 256x256 level with the L2 state of the real step (--cache-control none)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
    python scripts/profile_step.py eager > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
}
cap gn_fwd_ring_v2 'gn_fwd_grid_kernel' 0 2
cap gn_bwd_direct_v1 'gn_bwd_grid2_kernel' 0 3
ls -la gpurun_out/*.ncu-rep
