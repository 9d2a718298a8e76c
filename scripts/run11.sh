bash scripts/gpu_tests.sh tests/test_gpu_attention.py 2>&1 | grep -v "^$" | tail -30
cat > /tmp/gnmb_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = [sys.argv[0]]
import scripts.gn_microbench
PY
sed -i 's/^SHAPES = .*/SHAPES = [(65536, 256)]/' scripts/gn_microbench.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_.*_grid_kernel -s 4 -c 2 -f -o gpurun_out/gn_grid python scripts/gn_microbench.py > gpurun_out/ncu_gn_grid.log 2>&1; tail -3 gpurun_out/ncu_gn_grid.log
timeout 900 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-260
