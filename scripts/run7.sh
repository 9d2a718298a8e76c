python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for pdl in 1 0; do echo "== CGD_PDL=$pdl"; CGD_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['finite'])"; done
timeout 1200 python scripts/conv_sweep.py > gpurun_out/conv_sweep.txt 2>&1; grep -E "best|plan" gpurun_out/conv_sweep.txt
