"""GPU: the pair conv kernel with a split last wave (CGD_CONV_TAIL=1; csrc/conv_sched.cuh, TAIL instantiation of conv_tc2_kernel).
The schedule arithmetic is verified exhaustively on the host
(tests/test_conv_sched.py); device-validated in round 2, the path stays opt-in (no measured gain)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu  # device-validated in round 2 (gpurun call A: passed on a B200), no longer opt-in

CHILD = r'''
import sys, torch as th, torch.nn.functional as F
from clip_guided_diffusion_b200.plan import Plan, pack_conv
ok = True
for (NB, H, W, Cin, Cout, taps, has_r) in [(1, 256, 256, 256, 256, 9, True), (1, 256, 256, 128, 256, 9, False), (1, 160, 256, 64, 512, 1, True),
                                           (2, 128, 128, 256, 128, 9, False)]:
    th.manual_seed(0)
    k = 3 if taps == 9 else 1
    w = th.randn(Cout, Cin, k, k) * (taps * Cin) ** -0.5
    b = th.randn(Cout) * 0.1
    plan = Plan(conv_impl=3)
    cw = pack_conv(plan, w, b, need_bwd=False, name="w")
    x = plan.act(NB, H, W, Cin, "x")
    res = plan.act(NB, H, W, Cout, "res") if has_r else None
    y = plan.conv(x, cw, res=res, name="c")
    plan.finalize("cuda")
    xv = plan.view(x.buf, (NB, H, W, Cin)).normal_()
    rv = plan.view(res.buf, (NB, H, W, Cout)).normal_() if has_r else None
    plan.run(); th.cuda.synchronize()
    ref = F.conv2d(xv.float().permute(0, 3, 1, 2), w.cuda(), b.cuda(), padding=1 if taps == 9 else 0).permute(0, 2, 3, 1)
    if has_r: ref = ref + rv.float()
    got = plan.view(y.buf, (NB, H, W, Cout)).float()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(NB, H, W, Cin, Cout, taps, has_r, "err", err, "finite", bool(th.isfinite(got).all()))
    ok &= err < 3e-3 and bool(th.isfinite(got).all())
sys.exit(0 if ok else 1)
'''


def test_split_tail_conv_matches_reference():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=root, env=dict(os.environ, CGD_CONV_TAIL="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
