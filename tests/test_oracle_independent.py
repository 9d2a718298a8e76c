"""CPU: the two third-party restatements for which no independent implementation is installed (`oracle/diffusion.py`,
`oracle/unet.py`; SURVEY.md 8c: "parity unpinned") checked against derivations that do not share their code path:

* the Gaussian algebra of one sampling step re-derived in float64 from the process definition alone -- Bayes' rule on
  q(x_{t-1} | x_0) q(x_t | x_{t-1}) for the posterior, the learned-range variance interpolation of Improved-DDPM eq. 15, classifier
  guidance as a mean shift (ancestral) / score shift (DDIM, eta = 0) -- instead of the closed-form coefficient tables the oracle
  (and guided-diffusion) carry;
* the UNet's attention block against torch's own `scaled_dot_product_attention` / `nn.MultiheadAttention` kernels for both QKV
  layouts, the timestep embedding against the sinusoid written out per element, the up / down ResBlock resampling against
  `nn.Upsample` / `nn.AvgPool2d`, GroupNorm32 + scale-shift against a per-group mean / variance in float64.

This narrows, not removes, the "unpinned" status: the restatement still cannot be compared with the pinned packages themselves."""
import math

import numpy as np
import pytest
import torch as th
import torch.nn.functional as F

from oracle import diffusion as od
from oracle import unet as ou


# ----------------------------------------------------------------------------------------------------------------- diffusion
def _process(respacing):
    """base betas -> the respaced process, from the definition: abar'_i = abar_{S_i}, beta'_i = 1 - abar'_i / abar'_{i-1}"""
    d = od.create_gaussian_diffusion(1000, "linear", respacing)
    base_abar = np.cumprod(1.0 - np.linspace(1e-4, 2e-2, 1000, dtype=np.float64))  # linear schedule of Ho et al., T = 1000
    abar = base_abar[np.asarray(d.timestep_map)]
    abar_prev = np.concatenate([[1.0], abar[:-1]])
    return d, abar, abar_prev, 1.0 - abar / abar_prev


def _bayes_posterior(x0, xt, abar_prev, beta):
    """q(x_{t-1} | x_t, x_0) as a product of two Gaussians in x_{t-1}: prior N(sqrt(abar_prev) x0, 1 - abar_prev) and likelihood
    x_t ~ N(sqrt(1 - beta) x_{t-1}, beta)"""
    a = 1.0 - beta
    prec = 1.0 / (1.0 - abar_prev) + a / beta
    mean = (math.sqrt(abar_prev) * x0 / (1.0 - abar_prev) + math.sqrt(a) * xt / beta) / prec
    return mean, 1.0 / prec


class _FixedModel:
    """a 'network' with a fixed output: eps in channels 0-2, variance logits in 3-5"""

    def __init__(self, out):
        self.out = out

    def __call__(self, x, ts, **kw):
        return self.out.to(x.dtype) + 0.0 * x.sum()  # keeps the autograd graph of the *_with_grad samplers alive


@pytest.mark.parametrize("respacing,t", [("1000", 700), ("1000", 1), ("25", 12), ("250", 100)])
def test_ancestral_step_equals_first_principles(respacing, t):
    d, abar, abar_prev, beta = _process(respacing)
    assert np.allclose(d.betas, beta, rtol=1e-10, atol=0) and np.allclose(d.alphas_cumprod, abar, rtol=1e-12)
    g = th.Generator().manual_seed(3)
    xt = th.randn(2, 3, 8, 8, generator=g)
    out = th.randn(2, 6, 8, 8, generator=g)
    grad = th.randn(2, 3, 8, 8, generator=g) * 5
    th.manual_seed(17)
    res = d.p_sample_with_grad(_FixedModel(out), xt, th.full((2,), t), clip_denoised=False, cond_fn=lambda x, tt, o, **kw: grad)
    th.manual_seed(17)
    noise = th.randn_like(xt)

    X, E, V, G, N = (z.double().numpy() for z in (xt, out[:, :3], out[:, 3:], grad, noise))
    x0 = (X - math.sqrt(1 - abar[t]) * E) / math.sqrt(abar[t])  # x_t = sqrt(abar) x0 + sqrt(1 - abar) eps, solved for x0
    mean, var_lo = _bayes_posterior(x0, X, abar_prev[t], beta[t])
    frac = (V + 1) / 2
    var = np.exp(frac * math.log(beta[t]) + (1 - frac) * math.log(var_lo))  # Improved-DDPM eq. 15
    want = mean + var * G + np.sqrt(var) * N  # Dhariwal & Nichol, Algorithm 1: N(mu + Sigma g, Sigma)
    assert np.abs(res["pred_xstart"].double().numpy() - x0).max() < 2e-5 * np.abs(x0).max()
    assert np.abs(res["sample"].double().numpy() - want).max() < 2e-5 * np.abs(want).max()
    assert abs(d.posterior_variance[t] - var_lo) < 1e-12 * var_lo


@pytest.mark.parametrize("respacing,t", [("ddim250", 180), ("ddim250", 1), ("ddim50", 30)])
def test_ddim_step_equals_first_principles(respacing, t):
    d, abar, abar_prev, _ = _process(respacing)
    g = th.Generator().manual_seed(4)
    xt = th.randn(2, 3, 8, 8, generator=g)
    out = th.randn(2, 6, 8, 8, generator=g)
    grad = th.randn(2, 3, 8, 8, generator=g) * 5
    res = d.ddim_sample_with_grad(_FixedModel(out), xt, th.full((2,), t), clip_denoised=False, cond_fn=lambda x, tt, o, **kw: grad)
    X, E, G = (z.double().numpy() for z in (xt, out[:, :3], grad))
    x0 = (X - math.sqrt(1 - abar[t]) * E) / math.sqrt(abar[t])
    e_g = E - math.sqrt(1 - abar[t]) * G  # Dhariwal & Nichol, Algorithm 2: eps - sqrt(1 - abar) grad
    x0_g = (X - math.sqrt(1 - abar[t]) * e_g) / math.sqrt(abar[t])
    want = math.sqrt(abar_prev[t]) * x0_g + math.sqrt(1 - abar_prev[t]) * e_g  # Song et al. eq. 12 with sigma = 0
    assert np.abs(res["sample"].double().numpy() - want).max() < 2e-5 * np.abs(want).max()
    assert np.abs(res["pred_xstart"].double().numpy() - x0).max() < 2e-5 * np.abs(x0).max()  # the fork returns the UNGUIDED x0


def test_t0_adds_no_noise_and_q_sample_inverts():
    d, abar, _, _ = _process("25")
    g = th.Generator().manual_seed(5)
    xt, out = th.randn(1, 3, 4, 4, generator=g), th.randn(1, 6, 4, 4, generator=g)
    a = d.p_sample_with_grad(_FixedModel(out), xt, th.zeros(1, dtype=th.long), clip_denoised=False)
    b = d.p_sample_with_grad(_FixedModel(out), xt, th.zeros(1, dtype=th.long), clip_denoised=False)
    assert th.equal(a["sample"], b["sample"])  # two different noise draws, no effect at t = 0
    x0, eps = th.randn(1, 3, 4, 4, generator=g), th.randn(1, 3, 4, 4, generator=g)
    t = th.tensor([13])
    assert th.allclose(d._predict_xstart_from_eps(d.q_sample(x0, t, eps), t, eps), x0, atol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------- unet
@pytest.mark.parametrize("new_order", [False, True], ids=["legacy_qkv", "new_qkv"])
def test_attention_block_equals_torch_sdpa(new_order):
    th.manual_seed(0)
    C, nh, H = 64, 4, 6
    blk = ou.AttentionBlock(C, num_head_channels=C // nh, use_new_attention_order=new_order).eval()
    for p in blk.parameters():
        p.data.normal_(0, 0.2)
    x = th.randn(2, C, H, H)
    d = C // nh
    qkv = F.conv1d(F.group_norm(x.reshape(2, C, -1), 32, blk.norm.weight, blk.norm.bias, 1e-5), blk.qkv.weight, blk.qkv.bias)  # [2, 3C, T]
    if new_order:  # channels = [q (heads x d) | k | v]
        q, k, v = qkv.view(2, 3, nh, d, -1).unbind(1)
    else:  # channels = heads x [q | k | v]
        q, k, v = qkv.view(2, nh, 3, d, -1).unbind(2)
    o = F.scaled_dot_product_attention(q.transpose(-1, -2), k.transpose(-1, -2), v.transpose(-1, -2))  # [2, nh, T, d], scale 1 / sqrt(d)
    want = x + F.conv1d(o.transpose(-1, -2).reshape(2, C, -1), blk.proj_out.weight, blk.proj_out.bias).view_as(x)
    got = blk(x)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5


def test_new_order_attention_equals_nn_multihead_attention():
    """the new-order layout is nn.MultiheadAttention's packed in_proj layout: same weights, same result"""
    th.manual_seed(1)
    C, nh, T = 32, 2, 10
    blk = ou.AttentionBlock(C, num_head_channels=C // nh, use_new_attention_order=True).eval()
    for p in blk.parameters():
        p.data.normal_(0, 0.2)
    mha = th.nn.MultiheadAttention(C, nh, batch_first=True).eval()
    mha.in_proj_weight.data.copy_(blk.qkv.weight[:, :, 0])
    mha.in_proj_bias.data.copy_(blk.qkv.bias)
    mha.out_proj.weight.data.copy_(blk.proj_out.weight[:, :, 0])
    mha.out_proj.bias.data.copy_(blk.proj_out.bias)
    x = th.randn(3, C, T, 1)
    xn = F.group_norm(x.reshape(3, C, T), 32, blk.norm.weight, blk.norm.bias, 1e-5).transpose(1, 2)  # [B, T, C]
    want = x.reshape(3, C, T) + mha(xn, xn, xn, need_weights=False)[0].transpose(1, 2)
    assert float((blk(x).reshape(3, C, T) - want).abs().max() / want.abs().max()) < 1e-5


def test_timestep_embedding_per_element():
    t = th.tensor([0.0, 1.0, 417.0, 999.0])
    e = ou.timestep_embedding(t, 16).double().numpy()
    for n, tv in enumerate(t.tolist()):
        for i in range(8):
            w = 10000.0 ** (-i / 8)
            assert abs(e[n, i] - math.cos(tv * w)) < 2e-4 and abs(e[n, 8 + i] - math.sin(tv * w)) < 2e-4


@pytest.mark.parametrize("kind", ["plain", "up", "down"])
def test_resblock_equals_explicit_formula(kind):
    th.manual_seed(2)
    Cin, Cout, E = 32, 64, 48
    blk = ou.ResBlock(Cin, E, Cout, up=kind == "up", down=kind == "down").eval()
    for p in blk.parameters():
        p.data.normal_(0, 0.2)
    x, emb = th.randn(2, Cin, 8, 8), th.randn(2, E)

    def gn64(v, w, b):  # GroupNorm(32 groups, eps 1e-5) from its definition, in float64
        B, C, H, W = v.shape
        z = v.double().view(B, 32, -1)
        z = (z - z.mean(-1, keepdim=True)) / (z.var(-1, unbiased=False, keepdim=True) + 1e-5).sqrt()
        return z.view(B, C, H, W) * w.double().view(1, C, 1, 1) + b.double().view(1, C, 1, 1)

    silu = lambda v: v * th.sigmoid(v)
    resample = {"plain": lambda v: v, "up": th.nn.Upsample(scale_factor=2, mode="nearest"), "down": th.nn.AvgPool2d(2)}[kind]
    w = {k: v.double() for k, v in blk.state_dict().items()}
    h = resample(silu(gn64(x, w["in_layers.0.weight"], w["in_layers.0.bias"])))
    xs = resample(x.double())
    h = F.conv2d(h, w["in_layers.2.weight"], w["in_layers.2.bias"], padding=1)
    e = F.linear(silu(emb.double()), w["emb_layers.1.weight"], w["emb_layers.1.bias"])
    scale, shift = e[:, :Cout, None, None], e[:, Cout:, None, None]
    h = gn64(h, w["out_layers.0.weight"], w["out_layers.0.bias"]) * (1 + scale) + shift  # use_scale_shift_norm
    h = F.conv2d(silu(h), w["out_layers.3.weight"], w["out_layers.3.bias"], padding=1)
    want = F.conv2d(xs, w["skip_connection.weight"], w["skip_connection.bias"]) + h
    got = blk(x, emb).double()
    assert got.shape == want.shape and float((got - want).abs().max() / want.abs().max()) < 1e-5
