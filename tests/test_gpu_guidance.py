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
    assert out.dtype == th.float32 and np.allclose(out.cpu().numpy(), g["cut_out"], atol=1e-5)  # fp32 like the reference's pooled cutouts
    th.manual_seed(3)  # the reference's CPU-generator draw order
    out2 = MakeCutouts(32, 3)(src)
    assert np.allclose(out2.cpu().numpy(), g["cut_out"], atol=1e-5)
    up = MakeCutouts(56, 2)
    up.cached_coords = [tuple(int(v) for v in r) for r in g["cut_up_coords"]]
    assert np.allclose(up(T(g["cut_up_src"]), use_cache=True).cpu().numpy(), g["cut_up_out"], atol=1e-5)


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


def test_use_augs_cutout_kernels_vs_reference_goldens():
    """MakeCutouts(use_augs=True): the CUDA gather / scatter kernels (csrc/augs.cu) against tests/golden/augs_golden.npz -- outputs and
    autograd gradients of the REFERENCE's own module (cgd/modules.py + torchvision transforms, CPU) with the parameters and noise its
    seed produces.  fp16 patch output; nearest-neighbour sampling may pick the neighbouring pixel where a source coordinate lies
    within float round-off of .5, hence an L2 criterion next to the element-wise one."""
    import os
    from clip_guided_diffusion_b200.plan import Plan
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augs_golden.npz"))
    B, H, W, CS, CUTN = (int(v) for v in d["meta"])
    P, kpad = CS, 3 * CS * CS
    for s in d["seeds"]:
        s = int(s)
        plan = Plan()
        bx, bc, bp = plan.new(B * 3 * H * W, "f", "x"), plan.new(CUTN * 3, "i32", "coords"), plan.new(CUTN * 20, "f", "prm")
        bn = plan.new(CUTN * 4 * B * 3 * H * W, "f", "noise")
        bo, bd, bg = plan.new(CUTN * B * kpad, "h", "patches"), plan.new(CUTN * B * kpad, "h", "dpatches"), plan.new(B * 3 * H * W, "f", "dx")
        plan.emit("CUTOUTS_AUG_FWD", i=[B, H, W, CUTN, CS, P, kpad, min(H, W)], f=[0, 0, 0, 1, 1, 1], p=[(bx, 0), (bc, 0), (bo, 0), (bp, 0), (bn, 0)])
        plan.emit("FILL", i=[B * 3 * H * W], f=[0.0], p=[(bg, 0)])
        plan.emit("CUTOUTS_AUG_BWD", i=[B, H, W, CUTN, CS, P, kpad], f=[0, 0, 0, 1, 1, 1, 1.0], p=[(bd, 0), (bc, 0), (bg, 0), (bp, 0)])
        plan.finalize("cuda")
        plan.view(bx, (B, 3, H, W)).copy_(T(d[f"x_{s}"]) * 2 - 1)
        plan.view(bc, (CUTN, 3)).copy_(T(d[f"coords_{s}"]))
        plan.view(bp, (CUTN, 20)).copy_(T(d[f"prm_{s}"]))
        plan.view(bn).copy_(T(d[f"noise_{s}"]).flatten())
        plan.view(bd, (CUTN * B, 1, kpad)).copy_(T(d[f"cot_{s}"]).reshape(CUTN * B, 1, kpad))
        plan.run()
        th.cuda.synchronize()
        y = plan.view(bo, (CUTN * B, 3, CS, CS)).float().cpu()
        ref = th.from_numpy(d[f"y_{s}"])
        assert float((y - ref).norm() / ref.norm()) < 2e-3, (s, float((y - ref).norm() / ref.norm()))
        assert float(((y - ref).abs() > 5e-3).float().mean()) < 5e-3
        gx = plan.view(bg, (B, 3, H, W)).cpu()
        gref = th.from_numpy(d[f"gx_{s}"]) * 0.5  # the op differentiates through its own (x_in + 1) / 2
        assert float((gx - gref).norm() / gref.norm()) < 2e-2, (s, float((gx - gref).norm() / gref.norm()))


def test_make_cutouts_use_augs_surface():
    """the drop-in surface: MakeCutouts(cut_size, n, power, use_augs=True)(x) returns [cutn * B, 3, cut_size, cut_size] like the
    reference.  (On a GPU the reference takes its noise from the CUDA generator and its decisions from the CPU one; the goldens were
    made on the CPU where both interleave on one stream, so values are compared by the kernel test above, not here.)"""
    mk = MakeCutouts(24, 5, 1.0, use_augs=True)
    th.manual_seed(0)
    x = th.rand(2, 3, 48, 48, device="cuda")
    out = mk(x)
    assert out.shape == (10, 3, 24, 24) and th.isfinite(out).all()
    assert 0.2 < float(out.mean()) < 0.6 and float(out.min()) > -0.1 and float(out.max()) < 1.1  # [0, 1] images, zero fill, .01 noise
    th.manual_seed(0)
    x2 = th.rand(2, 3, 48, 48, device="cuda")
    assert th.equal(mk(x2), out)  # same seeds -> same windows, parameters and noise
