#!/bin/bash
# Round 2, call H: the final state as the driver will see it -- whole GPU suite, smoke(), both bench arms -- plus the records to commit
# (launch lists of one step and of the bench command, ncu --set full of every kernel family, the other BASELINE workloads).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
echo "=== pytest -m gpu (whole suite, one process)"
timeout 1700 python -m pytest tests/ -x -q -m gpu --durations=8 -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/r02_pytest_gpu_all_v2.log
echo "=== smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench, reference arm then own arm (driver order)"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tee gpurun_out/r02_bench_reference_v2.json | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_v3.json | cut -c1-600
tail -3 gpurun_out/bench.err
echo "=== launch lists: one eager step (warm), and the bench command itself"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v5_warm.csv python scripts/profile_step.py eager > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 900 --csv --log-file gpurun_out/r02_launches_bench_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
echo "=== ncu --set full, one launch per kernel family (final state)"
timeout 1200 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_families python scripts/profile_families.py > gpurun_out/ncu_families.log 2>&1
tail -2 gpurun_out/ncu_families.log
ncu -i /tmp/r02_families.ncu-rep --page raw --csv > gpurun_out/r02_families_v2_raw.csv 2>/dev/null; ls -la gpurun_out/r02_families_v2_raw.csv
echo "=== other BASELINE workloads (per-GPU shards)"
for w in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); t=d.get('torch_cuda_baseline',{}); print('$w', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],2), 'torch', round(t.get('value',0),2), 'vs_torch', round(d.get('vs_torch_cuda',0),2), 'step_tensor_frac', round(d['step_tensor_frac'],3))" | tee -a gpurun_out/r02_bench_workloads_v2.txt
done
du -sh gpurun_out
