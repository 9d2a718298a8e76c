#!/bin/bash
# ncu --set full captures of the dominant kernels of one eager cfg2 step (one invocation per kernel family)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
    python scripts/profile_step.py eager > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap conv3x3_256 conv_tc2_kernel 1 2
cap gn_fwd 'gn_(stats|apply)' 2 2
cap gn_bwd 'gn_bwd' 196 4
ls -la gpurun_out/*.ncu-rep
