"""fp32 restatement of guided-diffusion's sampler (oracle; test infrastructure).

Third-party (crowsonkb/guided-diffusion@fb47224 ``gaussian_diffusion.py`` / ``respace.py``), absent
from /root/reference; the reference reaches it at ``cgd/cgd.py:242-262`` (loops, kwargs) and reads
``num_timesteps`` / ``sqrt_one_minus_alphas_cumprod`` (``cgd/cgd.py:142,154,177,265``).  Restated
from SURVEY.md Appendix A.2 and call stacks 3.2 / 3.3: epsilon-prediction, LEARNED_RANGE variance,
the fork's ``*_with_grad`` variants, ``skip_timesteps`` / ``init_image`` / ``randomize_class``.
PARITY UNPINNED; structural pins = schedule constants in SURVEY.md Appendix E.
"""
from __future__ import annotations

import math

import numpy as np
import torch as th


def named_beta_schedule(name: str, T: int) -> np.ndarray:
    if name == "linear":
        scale = 1000.0 / T
        return np.linspace(scale * 1e-4, scale * 0.02, T, dtype=np.float64)
    if name == "cosine":
        f = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
        return np.array([min(1 - f((i + 1) / T) / f(i / T), 0.999) for i in range(T)], dtype=np.float64)
    raise NotImplementedError(name)


def space_timesteps(num_timesteps: int, section_counts) -> set:
    """'ddimN' -> fixed integer stride giving exactly N steps; 'a,b,c' / 'N' -> evenly spaced
    (rounded) indices per section."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {want} steps with an integer stride")
        section_counts = [int(s) for s in section_counts.split(",")]
    size, extra = divmod(num_timesteps, len(section_counts))
    start, out = 0, []
    for i, cnt in enumerate(section_counts):
        n = size + (1 if i < extra else 0)
        if n < cnt:
            raise ValueError(f"cannot divide section of {n} steps into {cnt}")
        frac = 1 if cnt <= 1 else (n - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            out.append(start + round(cur))
            cur += frac
        start += n
    return set(out)


def _extract(arr: np.ndarray, t: th.Tensor, shape) -> th.Tensor:
    res = th.from_numpy(arr).to(device=t.device)[t].float()
    while res.dim() < len(shape):
        res = res[..., None]
    return res.expand(shape)


class GaussianDiffusion:
    """EPSILON mean type, LEARNED_RANGE variance type (every checkpoint of the reference)."""

    def __init__(self, betas: np.ndarray, rescale_timesteps: bool = False):
        betas = np.asarray(betas, dtype=np.float64)
        self.betas = betas
        self.rescale_timesteps = rescale_timesteps
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        ac = np.cumprod(alphas)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - ac)

    # ---- closed forms
    def q_sample(self, x_start, t, noise):
        return (_extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + _extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def q_posterior_mean(self, x_start, x_t, t):
        return (_extract(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + _extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)

    def _predict_xstart_from_eps(self, x_t, t, eps):
        return (_extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - _extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return ((_extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - pred_xstart)
                / _extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape))

    def _scale_timesteps(self, t):
        return t.float() * (1000.0 / self.num_timesteps) if self.rescale_timesteps else t

    def _wrap(self, model):
        return model

    # ---- one model evaluation
    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        model_kwargs = model_kwargs or {}
        B, C = x.shape[:2]
        out = self._wrap(model)(x, self._scale_timesteps(t), **model_kwargs)
        assert out.shape == (B, 2 * C, *x.shape[2:])
        eps, v = th.split(out, C, dim=1)
        min_log = _extract(self.posterior_log_variance_clipped, t, x.shape)
        max_log = _extract(np.log(self.betas), t, x.shape)
        frac = (v + 1) / 2
        log_var = frac * max_log + (1 - frac) * min_log
        var = th.exp(log_var)
        x0 = self._predict_xstart_from_eps(x, t, eps)
        if denoised_fn is not None:
            x0 = denoised_fn(x0)
        if clip_denoised:
            x0 = x0.clamp(-1, 1)
        mean = self.q_posterior_mean(x0, x, t)
        return {"mean": mean, "variance": var, "log_variance": log_var, "pred_xstart": x0}

    # ---- conditioning (fork's *_with_grad variants: cond_fn receives the p_mean_variance dict)
    def condition_mean_with_grad(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        g = cond_fn(x, t, p_mean_var, **(model_kwargs or {}))
        return p_mean_var["mean"].float() + p_mean_var["variance"] * g.float()

    def condition_score_with_grad(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        abar = _extract(self.alphas_cumprod, t, x.shape)
        eps = self._predict_eps_from_xstart(x, t, p_mean_var["pred_xstart"])
        eps = eps - (1 - abar).sqrt() * cond_fn(x, t, p_mean_var, **(model_kwargs or {}))
        out = dict(p_mean_var)
        out["pred_xstart"] = self._predict_xstart_from_eps(x, t, eps)
        out["mean"] = self.q_posterior_mean(out["pred_xstart"], x, t)
        return out

    def p_sample_with_grad(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None):
        with th.enable_grad():
            x = x.detach().requires_grad_()
            out = self.p_mean_variance(model, x, t, clip_denoised, denoised_fn, model_kwargs)
            noise = th.randn_like(x)  # drawn BEFORE cond_fn
            if cond_fn is not None:
                out["mean"] = self.condition_mean_with_grad(cond_fn, out, x, t, model_kwargs)
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        sample = out["mean"] + nonzero * th.exp(0.5 * out["log_variance"]) * noise
        return {"sample": sample.detach(), "pred_xstart": out["pred_xstart"].detach()}

    def ddim_sample_with_grad(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None,
                              model_kwargs=None, eta=0.0):
        with th.enable_grad():
            x = x.detach().requires_grad_()
            out_orig = self.p_mean_variance(model, x, t, clip_denoised, denoised_fn, model_kwargs)
            out = out_orig
            if cond_fn is not None:
                out = self.condition_score_with_grad(cond_fn, out_orig, x, t, model_kwargs)
        out["pred_xstart"] = out["pred_xstart"].detach()
        eps = self._predict_eps_from_xstart(x, t, out["pred_xstart"])
        abar = _extract(self.alphas_cumprod, t, x.shape)
        abar_prev = _extract(self.alphas_cumprod_prev, t, x.shape)
        sigma = eta * th.sqrt((1 - abar_prev) / (1 - abar)) * th.sqrt(1 - abar / abar_prev)
        noise = th.randn_like(x)  # drawn AFTER cond_fn, consumed even when eta == 0
        mean_pred = out["pred_xstart"] * th.sqrt(abar_prev) + th.sqrt(1 - abar_prev - sigma ** 2) * eps
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        sample = mean_pred + nonzero * sigma * noise
        return {"sample": sample.detach(), "pred_xstart": out_orig["pred_xstart"].detach()}

    # ---- loops (call site cgd/cgd.py:250-262)
    def _loop(self, step_fn, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device,
              skip_timesteps, init_image, randomize_class, **extra):
        device = device or next(model.parameters()).device
        img = noise if noise is not None else th.randn(*shape, device=device)
        if skip_timesteps and init_image is None:
            init_image = th.zeros_like(img)
        indices = list(range(self.num_timesteps - skip_timesteps))[::-1]
        if init_image is not None:
            t0 = th.tensor([indices[0]] * shape[0], device=device, dtype=th.long)
            img = self.q_sample(init_image, t0, img)
        model_kwargs = dict(model_kwargs or {})
        for i in indices:
            t = th.tensor([i] * shape[0], device=device, dtype=th.long)
            if randomize_class and "y" in model_kwargs:
                model_kwargs["y"] = th.randint(0, model.num_classes, model_kwargs["y"].shape, device=device)
            with th.no_grad():
                out = step_fn(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                              cond_fn=cond_fn, model_kwargs=model_kwargs, **extra)
            yield out
            img = out["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                  cond_fn=None, model_kwargs=None, device=None, progress=False,
                                  skip_timesteps=0, init_image=None, randomize_class=False, cond_fn_with_grad=False):
        assert cond_fn_with_grad or cond_fn is None, "oracle restates only the *_with_grad path the reference uses"
        yield from self._loop(self.p_sample_with_grad, model, shape, noise, clip_denoised, denoised_fn, cond_fn,
                              model_kwargs, device, skip_timesteps, init_image, randomize_class)

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0,
                                     skip_timesteps=0, init_image=None, randomize_class=False, cond_fn_with_grad=False):
        assert cond_fn_with_grad or cond_fn is None
        yield from self._loop(self.ddim_sample_with_grad, model, shape, noise, clip_denoised, denoised_fn, cond_fn,
                              model_kwargs, device, skip_timesteps, init_image, randomize_class, eta=eta)


class _WrappedModel:
    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model, self.rescale, self.T0 = model, rescale_timesteps, original_num_steps
        self.timestep_map = timestep_map

    def __call__(self, x, ts, **kw):
        new_ts = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)[ts]
        if self.rescale:
            new_ts = new_ts.float() * (1000.0 / self.T0)
        return self.model(x, new_ts, **kw)


class SpacedDiffusion(GaussianDiffusion):
    """Keep only ``use_timesteps`` of a base schedule; betas recomputed from the kept alpha-bars."""

    def __init__(self, use_timesteps, betas, rescale_timesteps=False):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(betas)
        base_ac = np.cumprod(1.0 - np.asarray(betas, dtype=np.float64))
        last, new_betas, self.timestep_map = 1.0, [], []
        for i, ac in enumerate(base_ac):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        super().__init__(np.array(new_betas), rescale_timesteps=rescale_timesteps)

    def _wrap(self, model):
        return model if isinstance(model, _WrappedModel) else _WrappedModel(
            model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t  # scaling happens in the wrapper


def create_gaussian_diffusion(steps=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=False):
    betas = named_beta_schedule(noise_schedule, steps)
    return SpacedDiffusion(space_timesteps(steps, timestep_respacing or [steps]) if timestep_respacing
                           else set(range(steps)), betas, rescale_timesteps=rescale_timesteps)
