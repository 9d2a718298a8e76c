"""Generate golden vectors by running the reference's own importable hot-path modules.

Run in the BUILD container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/guidance_golden.npz (committed).  The modules imported are the live pieces of
the reference path: cgd/losses.py, cgd/modules.py (MakeCutouts), cgd/ResizeRight (tensor path).
cgd/cgd.py, cgd/clip_util.py and cgd/script_util.py cannot be imported (clip / lpips /
guided_diffusion are not installed and cannot be fetched), see SURVEY.md section 0.
"""
import os
import sys

import numpy as np
import torch as th

REF = "/root/reference"
sys.path.insert(0, REF)
from cgd import losses as ref_losses  # noqa: E402
from cgd.modules import MakeCutouts as RefMakeCutouts  # noqa: E402

out = {}

# --- losses: known answers + a larger seeded case with autograd gradients
th.manual_seed(0)
x = th.rand(1, 3)
y = th.rand(1, 3)
out["sph_small_x"], out["sph_small_y"] = x.numpy(), y.numpy()
out["sph_small"] = ref_losses.spherical_dist_loss(x, y).numpy()

th.manual_seed(0)
img = th.randn(2, 3, 8, 8) * 1.5
out["img_small"] = img.numpy()
out["range_small"] = ref_losses.range_loss(img).numpy()
out["tv_small"] = ref_losses.tv_loss(img).numpy()

g = th.Generator().manual_seed(7)
img = (th.randn(2, 3, 40, 48, generator=g) * 0.9).requires_grad_()
out["img_med"] = img.detach().numpy()
r = ref_losses.range_loss(img)
t = ref_losses.tv_loss(img)
out["range_med"], out["tv_med"] = r.detach().numpy(), t.detach().numpy()
out["range_med_grad"] = th.autograd.grad(r.sum(), img, retain_graph=True)[0].numpy()
out["tv_med_grad"] = th.autograd.grad(t.sum(), img)[0].numpy()

emb = th.randn(4, 2, 16, generator=g).requires_grad_()
tgt = th.randn(1, 16, generator=g)
d = ref_losses.spherical_dist_loss(emb.unsqueeze(0), tgt.unsqueeze(0))
out["sph_emb"], out["sph_tgt"], out["sph_med"] = emb.detach().numpy(), tgt.numpy(), d.detach().numpy()
out["sph_med_grad"] = th.autograd.grad(d.sum(), emb)[0].numpy()

# --- MakeCutouts: coordinate law under the CPU generator, pooling values and gradients
th.manual_seed(0)
mk = RefMakeCutouts(224, 4, 1.0)
out["coords_256_seed0"] = np.array(mk._generate_coords(256, 256, 4), dtype=np.int64)
th.manual_seed(0)
mk = RefMakeCutouts(224, 6, 0.5)
out["coords_512_pow05_seed0"] = np.array(mk._generate_coords(512, 512, 6), dtype=np.int64)
th.manual_seed(0)
out["raw_draws_seed0"] = np.array([float(th.rand([])), float(th.randint(0, 10, ())), float(th.randint(0, 10, ()))])

g = th.Generator().manual_seed(11)
src = th.rand(2, 3, 96, 96, generator=g).requires_grad_()
mk = RefMakeCutouts(32, 3, 1.0)
th.manual_seed(3)
cut = mk(src)
th.manual_seed(3)
out["cut_coords"] = np.array(mk._generate_coords(96, 96, 3), dtype=np.int64)
out["cut_src"] = src.detach().numpy()
out["cut_out"] = cut.detach().numpy()
wgt = th.randn(cut.shape, generator=g)
out["cut_wgt"] = wgt.numpy()
out["cut_grad"] = th.autograd.grad((cut * wgt).sum(), src)[0].numpy()

# up-sampling regime (64^2 checkpoints: S=64 -> cut_size 224)
src2 = th.rand(1, 3, 16, 16, generator=g)
mk = RefMakeCutouts(56, 2, 1.0)
th.manual_seed(5)
out["cut_up_out"] = mk(src2).numpy()
out["cut_up_src"] = src2.numpy()
th.manual_seed(5)
out["cut_up_coords"] = np.array(mk._generate_coords(16, 16, 2), dtype=np.int64)

# --- ResizeRight lanczos3 sample (named by north_star; dead at run time in the reference)
try:
    from cgd.ResizeRight import resize_right, interp_methods
    th.manual_seed(0)
    tt = th.rand(1, 1, 8, 8)
    out["rr_src"] = tt.numpy()
    out["rr_lanczos3_4x4"] = resize_right.resize(tt, out_shape=[4, 4], interp_method=interp_methods.lanczos3).numpy()
except Exception as e:  # pragma: no cover
    print("ResizeRight golden skipped:", e)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "guidance_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items()})
