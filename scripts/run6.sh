python scripts/gn_microbench.py 2>&1 | tail -22
for cs in 1 2 4; do echo "== GN_CS=$cs"; GN_CS=$cs python scripts/gn_microbench.py 2>&1 | grep fused | tail -10; done
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off -k regex:gn_fwd_fused_kernel -s 20 -c 2 -f -o gpurun_out/gn_fused python scripts/profile_step.py eager > gpurun_out/ncu_gn_fused.log 2>&1; tail -2 gpurun_out/ncu_gn_fused.log
