"""Per-launch efficiency report: rebuild the cfg2 op list on the CPU (shapes only, no weight values), pair every op with its
launch(es) of an ncu `gpu__time_duration` launch list (scripts/profile_step.py eager) and print FLOP/s per conv layer.

    python scripts/conv_layer_report.py profiles/r01_launches_cfg2_step_v3.csv
"""
import csv, os, re, sys, collections
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_guided_diffusion_b200 import plan as P

# shapes only: constants become plain allocations
P.Plan.const = lambda self, t, dt, name="": self.new(t.numel(), dt, name)
P.Plan.finalize = lambda self, device: self
import bench
from clip_guided_diffusion_b200 import weights as pw
pw.seeded_state_dict = lambda shapes, seed=0: {k: th.empty(v) for k, v in shapes.items()}
_orig_float = th.Tensor.float


def load_launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    out = []
    for r in rows[start:]:
        if len(r) <= mi:
            continue
        v = float(r[mi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        out.append((re.sub(r'\(.*', '', r[ki]), v))
    return out


def main():
    launches = load_launches(sys.argv[1])
    eng_args = {}
    from clip_guided_diffusion_b200 import guidance as pg
    pg.GuidedStepB200.set_targets = lambda *a, **k: None
    eng, diff, cond = bench.build_engine(th.device("cpu"), 0, 1)
    plan = eng.plan
    m = plan.marks
    segs = [("unet_emb", "unet_bwd"), ("pmv", "cond"), ("cut_fwd", "sph"), ("vit_fwd", "vit_bwd"), ("sph", "cut_bwd"), ("vit_bwd", "vit_end"),
            ("cut_bwd", "guide"), ("guide", "final"), ("unet_bwd", "unet_end"), ("final", "upd_anc_g"), ("upd_ddim_g", "upd_ddim")]
    ops = []
    for a, b in segs:
        ops += plan.ops[m[a]:m[b]]
    inv = {v: k for k, v in P.OP.items()}
    conv_l = [(n, t) for n, t in launches if 'conv_tc' in n or 'splitk' in n]
    ci = 0
    rows = []
    for op in ops:
        if inv[op.code] != "CONV":
            continue
        NB, H, W, Cin, Cout, npad, taps = op.i[:7]
        bn, splits = op.i[16], op.i[17]
        n, t = conv_l[ci]; ci += 1
        assert 'conv_tc' in n, n
        tr = 0.0
        if splits > 1:
            n2, tr = conv_l[ci]; ci += 1
            assert 'splitk' in n2, (n2, op.tag)
        fl = 2.0 * NB * H * W * npad * taps * Cin
        rows.append((op.tag, NB, H, W, Cin, Cout, taps, bn, splits, t, tr, fl))
    assert ci == len(conv_l), (ci, len(conv_l))
    tot_t = sum(r[9] + r[10] for r in rows)
    print(f"{'tag':46s} {'NBxHxW':>12s} {'Cin':>5s} {'Cout':>5s} t  BN sp {'us':>8s} {'red us':>7s} {'TF/s':>7s}")
    groups = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for tag, NB, H, W, Cin, Cout, taps, bn, sp, t, tr, fl in rows:
        key = (NB, H, W, Cin, Cout, taps, bn, sp)
        g = groups[key]; g[0] += 1; g[1] += t + tr; g[2] += fl
        if len(sys.argv) > 2:
            print(f"{tag[:46]:46s} {NB}x{H}x{W:>4d} {Cin:5d} {Cout:5d} {taps} {bn:3d} {sp:2d} {t:8.1f} {tr:7.1f} {fl / (t + tr) / 1e6:7.1f}")
    print("--- grouped by shape (count, total us, share of conv time, TF/s)")
    for key, (c, t, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        NB, H, W, Cin, Cout, taps, bn, sp = key
        print(f"{NB:3d}x{H:4d}x{W:4d} Cin {Cin:5d} Cout {Cout:5d} taps {taps} BN {bn:3d} sp {sp:2d}: n={c:3d} {t:8.1f} us {100 * t / tot_t:5.1f}% {fl / t / 1e6:7.1f} TF/s")
    print(f"conv total {tot_t:.1f} us, {sum(r[11] for r in rows) / 1e12:.3f} TFLOP (padded), {sum(r[11] for r in rows) / tot_t / 1e6:.1f} TF/s")


main()
