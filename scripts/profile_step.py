"""One guided step (cfg2, or the bench workload named by CGD_PROFILE_WORKLOAD) inside a cudaProfilerStart/Stop range (for
`ncu --profile-from-start off`)."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
if os.environ.get("CGD_PROFILE_WORKLOAD"):
    bench.CFG = bench.WORKLOADS[os.environ["CGD_PROFILE_WORKLOAD"]]
sampler = "ddim" if bench.CFG["respacing"].startswith("ddim") else "ancestral"
th.cuda.set_device(0)
eng, diff, cond = bench.build_engine(th.device("cuda", 0), 0, 1)
eng.use_graph = mode == "graph"
th.manual_seed(0)
img = eng.draw_initial_noise()
i = diff.num_timesteps - 1
for _ in range(2):
    img = eng.fused_step(diff, sampler, i, img, eng.draw_classes(), cond, 0.0)["sample"]
    cond.step_done()
    i -= 1
th.cuda.synchronize()
th.cuda.profiler.start()
img = eng.fused_step(diff, sampler, i, img, eng.draw_classes(), cond, 0.0)["sample"]
th.cuda.synchronize()
th.cuda.profiler.stop()
print("profiled one step; finite:", bool(th.isfinite(img).all()))
