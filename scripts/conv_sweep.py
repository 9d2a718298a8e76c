"""Sweep (BN, split-K) for the latency-bound conv / GEMM shapes of the step; each configuration is timed as a CUDA graph of
back-to-back launches over rotating buffers (like the real step: no CPU launch gaps).  python scripts/conv_sweep.py"""
import os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_guided_diffusion_b200 import plan as P

SHAPES = [  # NB, H, W, Cin, Cout, taps
    (1, 1, 800, 768, 768, 1), (1, 1, 800, 768, 2304, 1), (1, 1, 800, 768, 3072, 1), (1, 1, 800, 3072, 768, 1),
    (1, 8, 8, 1024, 1024, 9), (1, 8, 8, 1024, 1024, 1), (1, 16, 16, 1024, 1024, 9), (1, 16, 16, 1024, 1024, 1),
    (1, 32, 32, 512, 512, 9), (1, 32, 32, 512, 512, 1), (1, 32, 32, 1024, 512, 9), (1, 64, 64, 512, 512, 9),
    (1, 64, 64, 256, 256, 9), (1, 128, 128, 256, 256, 9),
]
dev = th.device("cuda", 0)
ORIG = P.pick_splits
only = sys.argv[1] if len(sys.argv) > 1 else None
for (NB, H, W, Cin, Cout, taps) in SHAPES:
    kblocks = taps * Cin // 64
    best = None
    for bn in (64, 128, 192, 256):
        if P._npad(Cout) % bn:
            continue
        for sp in (0, 1, 2, 3, 4, 6, 8, 12, 16, 24):  # 0 = the plan's own choice
            if sp > kblocks or (sp > 1 and kblocks // sp < 2):
                continue
            P.pick_bn = lambda npad, m_tiles, kblocks=0, bn=bn: bn
            if sp:
                P.pick_splits = lambda m_tiles, n_tiles, kb, npad, ws_cap_bytes=0, sp=sp: sp
            else:
                import importlib
                P.pick_splits = ORIG if 'ORIG' in globals() else P.pick_splits
            th.manual_seed(0)
            plan = P.Plan()
            k = 3 if taps == 9 else 1
            cw = P.pack_conv(plan, th.randn(Cout, Cin, k, k) * (taps * Cin) ** -0.5, th.zeros(Cout), need_bwd=False, name="w")
            nbuf = 8
            for _ in range(nbuf):
                x = plan.act(NB, H, W, Cin, "x")
                plan.conv(x, cw, name="c")
            plan.finalize(dev)
            for b in plan.bufs:
                if b.name == "x":
                    plan.view(b).normal_()
            plan.run(); plan.run()
            th.cuda.synchronize()
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                for _ in range(4):
                    plan.run()
            g.replay(); th.cuda.synchronize()
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record(); th.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3 / (5 * 4 * nbuf)
            op = [o for o in plan.ops if o.code == 1][0]
            tag = f"BN{op.i[16]:3d} sp{op.i[17]:2d}"
            print(f"{NB}x{H}x{W} {Cin}->{Cout} t{taps} {tag}{' (plan)' if sp == 0 else '':7s}: {t * 1e6:7.1f} us", flush=True)
            if best is None or t < best[0]:
                best = (t, tag)
            del plan, g
    print(f"  -> best {best[1]} {best[0] * 1e6:.1f} us", flush=True)
