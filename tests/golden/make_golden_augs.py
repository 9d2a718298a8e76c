"""Golden vectors for MakeCutouts(use_augs=True) produced by the REFERENCE's own module (cgd/modules.py:12-24, 50-66) with
torchvision's transforms, on the CPU.

Run in the BUILD container only (needs /root/reference):
    python tests/golden/make_golden_augs.py
Writes tests/golden/augs_golden.npz (committed).  Several seeds so that every branch (flip / no flip, perspective on / off,
grayscale on / off) occurs.  Stored per case: the input, the seed, the reference's output, its autograd gradient for a fixed
cotangent, the torch RNG state after the call -- and, re-derived by clip_guided_diffusion_b200/augs.py from the same seed, the
cutout windows, the 20 aug parameters per cutout and the four noise fields, so that the GPU kernels (no /root/reference on the GPU
box) can be checked against what the reference computed.  tests/test_oracle.py re-derives them again and checks (a) the draw order
(final RNG state identical to the reference's) and (b) oracle.apply_augs == the reference, bit for bit.
"""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")
from cgd.modules import MakeCutouts as RefMakeCutouts  # noqa: E402

from clip_guided_diffusion_b200 import augs  # noqa: E402

B, H, W, CS, CUTN = 2, 48, 48, 24, 4
out = {"meta": np.array([B, H, W, CS, CUTN])}
seeds = [0, 1, 5]
flags = []
for s in seeds:
    g = th.Generator().manual_seed(1000 + s)
    x = th.rand(B, 3, H, W, generator=g)
    cot = th.randn(CUTN * B, 3, CS, CS, generator=g)
    ref = RefMakeCutouts(CS, CUTN, 1.0, use_augs=True)
    th.manual_seed(s)
    xr = x.clone().requires_grad_()
    y = ref(xr)
    state = th.get_rng_state()
    (gx,) = th.autograd.grad((y * cot).sum(), xr)
    # the same stream through this repo's host code
    th.manual_seed(s)
    coords = ref._generate_coords(H, W, CUTN)  # sic: (H, W) as (side_x, side_y) like forward() does
    noise = th.zeros(CUTN, 4, B, 3, min(H, W), min(H, W))
    prm = augs.draw_aug_params(coords, B, H, W, noise_device="cpu", noise_out=noise)
    assert th.equal(th.get_rng_state(), state), f"seed {s}: draw order differs from the reference's"
    out[f"x_{s}"], out[f"cot_{s}"], out[f"y_{s}"], out[f"gx_{s}"] = x.numpy(), cot.numpy(), y.detach().numpy(), gx.numpy()
    out[f"coords_{s}"], out[f"prm_{s}"], out[f"noise_{s}"] = np.array(coords, dtype=np.int32), prm.numpy(), noise.numpy().astype(np.float32)
    out[f"rng_{s}"] = state.numpy()
    flags.append(prm[:, [0, 7, 16]].tolist())
out["seeds"] = np.array(seeds)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "augs_golden.npz"), **out)
print("flip / perspective / gray per cutout:", flags)
