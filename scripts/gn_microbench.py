"""GroupNorm kernels alone: fused (cluster) vs two-pass, forward and backward, per UNet shape; CUDA events, eager launches with
rotating buffers.  Usage: python scripts/gn_microbench.py"""
import os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_guided_diffusion_b200 import plan as P

SHAPES = [(64, 1024), (256, 1024), (1024, 512), (1024, 1024), (4096, 512), (4096, 256), (16384, 256), (16384, 512), (65536, 256), (65536, 512)]
dev = th.device("cuda", 0)
cs_override = int(os.environ.get("GN_CS", "0"))
if cs_override:
    orig = P.gn_fused_cluster
    P.gn_fused_cluster = lambda N, HW, C, maxv: (cs_override if orig(N, HW, C, maxv) and -(-HW // cs_override) <= maxv * (512 // max(2, C // 256)) else orig(N, HW, C, maxv))
for HW, C in SHAPES:
    for mode in ('auto', 'grid', 'twopass'):
        if os.environ.get('GN_ONLY') and mode != os.environ['GN_ONLY']:
            continue
        plan = P.Plan()
        plan.fused_gn = mode == 'auto'
        plan.grid_gn = mode != 'twopass'
        gb, bb = plan.const(th.ones(C), "f", "g"), plan.const(th.zeros(C), "f", "b")
        eb = plan.const(th.zeros(2 * C), "f", "e")
        nbuf = max(2, min(8, int(200e6 // (HW * C * 8)) + 1))
        for _ in range(nbuf):
            x = plan.act(1, 1, HW, C, "x")
            y = plan.group_norm(x, gb, bb, emb=(eb, 0), silu=True, name="gn")
            dy = plan.act(1, 1, HW, C, "dy")
            plan._grads[y.key()] = dy
        plan.mark("bwd")
        plan.backward()
        plan.mark("end")
        plan.finalize(dev)
        for b in plan.bufs:
            if b.name in ("x", "dy"):
                plan.view(b).normal_()
        res = []
        for a, bm in (("fwd", (0, plan.marks["bwd"])), ("bwd", (plan.marks["bwd"], plan.marks["end"]))):
            n_ops = bm[1] - bm[0]
            for _ in range(2):
                plan.run(bm[0], n_ops)
            th.cuda.synchronize()
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            rounds = max(1, 32 // nbuf)
            e0.record()
            for _ in range(rounds):
                plan.run(bm[0], n_ops)
            e1.record()
            th.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3 / (rounds * nbuf)
            byts = HW * C * (4 if a == "fwd" else 6)
            res.append(f"{a} {t * 1e6:7.1f} us {byts / t / 1e9:7.0f} GB/s")
        kinds = sorted({o.code for o in plan.ops})
        cs = [o.i[5] for o in plan.ops if o.code == 33][:1] + [o.i[6] for o in plan.ops if o.code == 34][:1]
        print(f"HW {HW:6d} C {C:5d} {'fused' if 33 in kinds else ('grid' if 35 in kinds else 'twopass'):8s} CS{cs}: " + " | ".join(res), flush=True)
        del plan
