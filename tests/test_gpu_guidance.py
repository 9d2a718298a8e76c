"""GPU: the stand-alone C-ABI guidance operators against the golden vectors generated from the reference's own modules
(tests/golden/guidance_golden.npz: cgd/losses.py, cgd/modules.py run in the build container)."""
import ctypes

import numpy as np
import pytest
import torch as th

from clip_guided_diffusion_b200 import _lib, losses as pl
from clip_guided_diffusion_b200.guidance import MakeCutouts
from tests.plan_interp import Interp

pytestmark = pytest.mark.gpu


def T(a):
    return th.from_numpy(np.asarray(a)).cuda()


def test_losses_golden(golden):
    g = golden
    assert np.allclose(pl.range_loss(T(g["img_small"])).cpu().numpy(), g["range_small"], rtol=1e-5, atol=1e-7)
    assert np.allclose(pl.tv_loss(T(g["img_small"])).cpu().numpy(), g["tv_small"], rtol=1e-5)
    assert np.allclose(pl.range_loss(T(g["img_med"])).cpu().numpy(), g["range_med"], rtol=1e-5, atol=1e-7)
    assert np.allclose(pl.tv_loss(T(g["img_med"])).cpu().numpy(), g["tv_med"], rtol=1e-5)
    d = pl.spherical_dist_loss(T(g["sph_small_x"]), T(g["sph_small_y"]))
    assert np.allclose(d.cpu().numpy(), g["sph_small"], rtol=1e-5)
    d = pl.spherical_dist_loss(T(g["sph_emb"]).unsqueeze(0), T(g["sph_tgt"]).unsqueeze(0))
    assert np.allclose(d.cpu().numpy(), g["sph_med"], rtol=2e-5)


def test_loss_gradients_golden(golden):
    """analytic d(tv)/dx and d(range)/dx from the fused kernel vs autograd of the reference's losses"""
    g = golden
    x = T(g["img_med"])
    _, dxd = pl._guide(x)  # tv_scale = range_scale = 1, fac = 0: dx_direct = d(sum tv)/dx + d(sum range)/dx
    ref = g["tv_med_grad"] + g["range_med_grad"]
    assert np.allclose(dxd.cpu().numpy(), ref, rtol=1e-4, atol=1e-7)


def test_cutouts_golden(golden):
    g = golden
    src = T(g["cut_src"])
    coords = [tuple(int(v) for v in r) for r in g["cut_coords"]]
    mk = MakeCutouts(32, 3)
    mk.cached_coords = coords
    out = mk(src, use_cache=True)
    assert out.shape == (6, 3, 32, 32)
    assert np.allclose(out.cpu().numpy(), g["cut_out"], atol=1e-3)  # fp16 patch output
    th.manual_seed(3)  # the reference's CPU-generator draw order
    out2 = MakeCutouts(32, 3)(src)
    assert np.allclose(out2.cpu().numpy(), g["cut_out"], atol=1e-3)
    up = MakeCutouts(56, 2)
    up.cached_coords = [tuple(int(v) for v in r) for r in g["cut_up_coords"]]
    assert np.allclose(up(T(g["cut_up_src"]), use_cache=True).cpu().numpy(), g["cut_up_out"], atol=1e-3)


def test_cutouts_backward_golden(golden):
    g = golden
    coords = th.tensor(g["cut_coords"], dtype=th.int32).cuda()
    wgt = T(g["cut_wgt"])  # upstream gradient d/d(cutouts) [6,3,32,32]; kernel sees d/d(normalised patches)
    n, cs = 6, 32
    dp = Interp._patchify(wgt.float().cpu(), cs, 3 * cs * cs).half().cuda().contiguous()
    dx = th.empty(2, 3, 96, 96, device="cuda")
    std = (ctypes.c_float * 3)(0.5, 0.5, 0.5)  # (0.5/std) == 1: plain pooled-cutout gradient
    rc = _lib.load().cgd_cutouts_bwd(ctypes.c_void_p(dp.data_ptr()), ctypes.c_void_p(coords.data_ptr()), ctypes.c_void_p(dx.data_ptr()),
                                     ctypes.c_int64(2), ctypes.c_int64(96), ctypes.c_int64(96), ctypes.c_int64(3), ctypes.c_int64(cs),
                                     ctypes.c_int64(cs), ctypes.c_int64(3 * cs * cs), std, ctypes.c_float(1.0),
                                     ctypes.c_void_p(th.cuda.current_stream().cuda_stream))
    _lib.check(rc)
    assert np.allclose(dx.cpu().numpy(), g["cut_grad"], atol=2e-3, rtol=2e-3)


def test_sharded_rms_clamp_ops_equal_fused():
    """FINAL_GRAD(partial sums only) + MAG_CLAMP -- the sharded-batch form of the whole-batch RMS clamp (cgd/cgd.py:229-232) --
    equals the two-launch FINAL_GRAD when the 'whole batch' is the local one, and scales with the global element count."""
    from clip_guided_diffusion_b200.plan import Plan
    B, HW = 2, 48 * 48
    n = B * 3 * HW
    outs = {}
    for name in ("fused", "split", "split_2x"):
        plan = Plan()
        dxd, dxu, g, ws = plan.new(n, "f", "dxd"), plan.new(n, "f", "dxu"), plan.new(n, "f", "g"), plan.new(128, "f", "ws")
        plan.emit("FINAL_GRAD", flags=1 if name == "fused" else 5, i=[B, HW], f=[0.5, 0.05], p=[(dxd, 0), (dxu, 0), (g, 0), (ws, 0)])
        if name != "fused":
            plan.emit("MAG_CLAMP", i=[n, n if name == "split" else 2 * n], f=[0.05], p=[(g, 0), (ws, 0)])
        plan.finalize("cuda")
        gen = th.Generator().manual_seed(1)
        plan.view(dxd).copy_(th.randn(n, generator=gen))
        plan.view(dxu).copy_(th.randn(n, generator=gen))
        plan.run()
        th.cuda.synchronize()
        outs[name] = plan.view(g).float().cpu().clone()
    assert abs(float(outs["fused"].square().mean().sqrt()) - 0.05) < 1e-5
    assert th.allclose(outs["split"], outs["fused"], rtol=1e-6, atol=1e-9)
    # twice the element count with the same sum of squares: the RMS halves by sqrt(2) but still exceeds the cap -> 0.05 * sqrt(2) locally
    assert abs(float(outs["split_2x"].square().mean().sqrt()) - 0.05 * 2 ** 0.5) < 1e-5
