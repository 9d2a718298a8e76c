"""``range_loss`` / ``spherical_dist_loss`` / ``tv_loss`` with the reference's signatures (cgd/losses.py:5-22), evaluated by
the CUDA guidance kernels (forward values; inside the sampling step the same kernels also emit the analytic gradients, so
no autograd graph exists on this path)."""
from __future__ import annotations

import ctypes

import torch as th

from . import _lib
from ._lib import SC


def _stream():
    return ctypes.c_void_p(th.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _guide(x: th.Tensor):
    if not x.is_cuda:
        raise _lib.CgdError("cgd losses run on the GPU only (no CPU fallback)")
    B, C, H, W = x.shape
    assert C == 3, "losses are defined on RGB images [B,3,H,W]"
    x = x.detach().float().contiguous()
    sc = th.zeros(SC["COUNT"], device=x.device)
    sc[SC["SQRT_RECIP_AC"]] = 1.0
    sc[SC["ONE_MINUS_FAC"]] = 1.0
    seed = th.empty(B * H * W * 8, dtype=th.float16, device=x.device)
    dxd = th.empty_like(x)
    loss = th.zeros(3 * B, device=x.device)
    lib = _lib.load()
    rc = lib.cgd_guidance_losses_fwd_bwd(_p(x), _p(x), None, _p(sc), _p(seed), _p(dxd), _p(loss), ctypes.c_int64(B), ctypes.c_int64(H),
                                         ctypes.c_int64(W), ctypes.c_int64(8), ctypes.c_float(1.0), ctypes.c_float(1.0), ctypes.c_float(0.0),
                                         ctypes.c_float(1.0), _stream())
    _lib.check(rc, "cgd_guidance_losses_fwd_bwd")
    return loss.view(3, B), dxd


def range_loss(input: th.Tensor) -> th.Tensor:
    return _guide(input)[0][1]


def tv_loss(input: th.Tensor) -> th.Tensor:
    return _guide(input)[0][0]


def spherical_dist_loss(x: th.Tensor, y: th.Tensor) -> th.Tensor:
    """x [..., D] against one target vector y [..., 1, D] / [D] (the broadcast the reference's call site produces for a single
    prompt, cgd/cgd.py:196-198)."""
    if not x.is_cuda:
        raise _lib.CgdError("cgd losses run on the GPU only (no CPU fallback)")
    D = x.shape[-1]
    if y.numel() != D:
        raise NotImplementedError("spherical_dist_loss kernel entry point takes one target vector; batched targets go through the engine")
    xs = x.detach().float().reshape(-1, D).contiguous()
    n = xs.shape[0]
    ys = y.detach().float().reshape(1, D).contiguous()
    w = th.ones(1, device=x.device)
    d_emb = th.empty_like(xs)
    loss = th.empty(n, device=x.device)
    rc = _lib.load().cgd_spherical_fwd_bwd(_p(xs), _p(ys), _p(w), _p(d_emb), _p(loss), ctypes.c_int64(1), ctypes.c_int64(n), ctypes.c_int64(1),
                                           ctypes.c_int64(D), ctypes.c_float(1.0), ctypes.c_float(1.0), _stream())
    _lib.check(rc, "cgd_spherical_fwd_bwd")
    return loss.view(th.broadcast_shapes(x.shape[:-1], y.shape[:-1]) if y.dim() > 1 else x.shape[:-1])
