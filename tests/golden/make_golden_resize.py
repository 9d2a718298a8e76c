"""Golden vectors for the ResizeRight cutout mode, produced by the REFERENCE's own vendored resampler (tensor path of
cgd/ResizeRight/resize_right.py:31-122 with interp_methods.lanczos3, antialiasing=True -- the call encode_image_prompt makes at
cgd/clip_util.py:95-97).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_resize.py        # writes tests/golden/resize_right_golden.npz (committed)
"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, "/root/reference")
from cgd.ResizeRight import interp_methods, resize_right  # noqa: E402

out = {}
g = th.Generator().manual_seed(11)
# square crops S -> cs, the sizes MakeCutouts produces: down-scaling (256^2 / 512^2 images -> 224), identity, up-scaling (64 -> 224)
# (small sizes with the same scale factors as the real ones keep the fixture ~1 MB; one case at the real size)
for name, S, cs in (("down_237_224", 237, 224), ("down_64_56", 64, 56), ("down_60_32", 60, 32), ("down_73_32", 73, 32),
                    ("same_32", 32, 32), ("up_16_56", 16, 56), ("up_20_32", 20, 32)):
    C = 3 if S <= 100 else 1
    x = (th.rand(1, C, S, S, generator=g) * 2 - 1).requires_grad_()
    y = resize_right.resize(x, out_shape=[cs, cs], interp_method=interp_methods.lanczos3, support_sz=None, antialiasing=True, by_convs=False)
    wgt = th.randn(y.shape, generator=g)
    (gx,) = th.autograd.grad((y * wgt).sum(), x)
    keep = slice(None)
    out[name + "_x"] = x.detach().numpy()[:, keep]
    out[name + "_y"] = y.detach().numpy()[:, keep]
    out[name + "_w"] = wgt.numpy()[:, keep]
    out[name + "_g"] = gx.numpy()[:, keep]
    out[name + "_shape"] = np.array([S, cs, C])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_right_golden.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith("_y")})
