"""Restatement of the reference's vendored ResizeRight resampler for the case the cutout mode uses (oracle; test
infrastructure): 2-D tensor [..., S, S] -> [..., O, O], lanczos3, antialiasing, pad_mode='constant', by_convs=False.
Follows cgd/ResizeRight/resize_right.py:31-122 (resize), :126-137 (get_projected_grid), :140-150 (get_field_of_view),
:153-164 (calc_pad_sz), :202-213 (get_weights), :216-247 (apply_weights), :341-353 (apply_antialiasing_if_needed) and
cgd/ResizeRight/interp_methods.py:51-55 (lanczos3).  PINNED: tests/test_oracle.py checks it against
tests/golden/resize_right_golden.npz, which tests/golden/make_golden_resize.py produced by running the reference's own module."""
from __future__ import annotations

from math import ceil, pi

import torch as th
import torch.nn.functional as F

EPS = float(th.finfo(th.float32).eps)


def lanczos3(x):
    return ((th.sin(pi * x) * th.sin(pi * x / 3) + EPS) / ((pi ** 2 * x ** 2 / 3) + EPS)) * (abs(x) < 3).to(x.dtype)


def _resize_dim(x, dim, out_sz):
    in_sz = x.shape[dim]
    scale = float(out_sz / in_sz)
    if scale == 1.0:
        return x
    grid = th.arange(out_sz) / scale + (in_sz - 1) / 2 - (out_sz - 1) / (2 * scale)
    if scale < 1.0:
        support = 6 / scale
        kern = lambda a: scale * lanczos3(scale * a)  # noqa: E731
    else:
        support, kern = 6, lanczos3
    left = (grid - support / 2 - EPS).ceil().long()
    fov = left[:, None] + th.arange(ceil(support - EPS))
    pad = [-int(fov[0, 0]), int(fov[-1, -1]) - in_sz + 1]
    fov = fov + pad[0]
    grid = grid + pad[0]
    w = kern(grid[:, None] - fov)
    sw = w.sum(1, keepdim=True)
    sw[sw == 0] = 1
    w = w / sw
    t = x.transpose(dim, 0)
    t = F.pad(t.transpose(0, -1), pad).transpose(0, -1)  # constant (zero) padding; negative = crop
    nb = t[fov]
    out = (nb * w.reshape(*w.shape, *([1] * (x.ndim - 1)))).sum(1)
    return out.transpose(0, dim)


def resize_lanczos3(x, out_hw):
    """x [..., H, W] -> [..., out_hw[0], out_hw[1]] (dimensions processed in order of increasing scale like the reference)"""
    dims = sorted([(x.ndim - 2, out_hw[0]), (x.ndim - 1, out_hw[1])], key=lambda d: d[1] / x.shape[d[0]])
    for dim, o in dims:
        x = _resize_dim(x, dim, o)
    return x
