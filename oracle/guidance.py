"""fp32 restatement of the reference's in-tree guidance code (oracle; test infrastructure).

Follows, and is pinned by golden vectors generated from, the importable reference modules:
``cgd/losses.py:5-22``, ``cgd/modules.py:5-66``, ``cgd/clip_util.py:45`` and the ``cond_fn`` closure
``cgd/cgd.py:151-239`` (which cannot itself be imported: its module imports clip/lpips).
Goldens: ``tests/golden/guidance_golden.npz`` made by ``tests/golden/make_golden.py``; the ``cond_fn`` restatement
(``OracleCondFn`` / ``guidance_loss`` / ``magnitude_clamp``) is pinned on ``tests/golden/cond_fn_golden.npz``: what the reference's
own closure returned when its unmodified source was cut out of ``cgd/cgd.py`` with ``ast`` and executed
(``tests/golden/make_golden_cond_fn.py``, ``tests/test_cond_fn_reference.py``: 1e-5 incl. the whole-batch RMS clamp, the
saturation term, several weighted prompts, a non-square image).
"""
from __future__ import annotations

import torch as th
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # cgd/clip_util.py:45
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def range_loss(x):  # cgd/losses.py:5-7
    return (x - x.clamp(-1, 1)).pow(2).mean([1, 2, 3])


def spherical_dist_loss(x, y):  # cgd/losses.py:10-14
    x = F.normalize(x, dim=-1)
    y = F.normalize(y, dim=-1)
    return (x - y).norm(dim=-1).div(2).arcsin().pow(2).mul(2)


def tv_loss(x):  # cgd/losses.py:17-22
    x = F.pad(x, (0, 1, 0, 1), "replicate")
    dx = x[..., :-1, 1:] - x[..., :-1, :-1]
    dy = x[..., 1:, :-1] - x[..., :-1, :-1]
    return (dx ** 2 + dy ** 2).mean([1, 2, 3])


def clip_normalize(x):  # cgd/clip_util.py:45 (torchvision Normalize)
    mean = th.tensor(CLIP_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    std = th.tensor(CLIP_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


def apply_augs(cut, prm, noise=None):
    """The reference's ``self.augs(cutout)`` (cgd/modules.py:12-24, 62) with EXPLICIT randomness: ``prm`` = the 20 numbers
    clip_guided_diffusion_b200/augs.py draws per cutout in torchvision's order (flip | inverse affine matrix | perspective on |
    coefficients | grayscale), ``noise`` = the four N(0, .01^2) fields [4, B, 3, S, S] (None: none).  The image operations are
    torchvision's own functional kernels -- the ones RandomAffine / RandomPerspective / RandomGrayscale call -- so this IS the
    reference pipeline, only with the random draws lifted out (pinned against cgd.modules.MakeCutouts(use_augs=True) itself in
    tests/test_oracle.py)."""
    import torchvision.transforms._functional_tensor as FT
    prm = [float(v) for v in prm]
    fill = [0.0, 0.0, 0.0]

    def nz(k, x):
        return x if noise is None else x + noise[k]

    x = cut.flip(-1) if prm[0] else cut
    x = nz(0, x)
    x = FT.affine(x, matrix=prm[1:7], interpolation="nearest", fill=fill)
    x = nz(1, x)
    if prm[7]:
        x = FT.perspective(x, prm[8:16], interpolation="bilinear", fill=fill)
    x = nz(2, x)
    if prm[16]:
        x = FT.rgb_to_grayscale(x, num_output_channels=3)
    return nz(3, x)


class MakeCutouts(th.nn.Module):
    """cgd/modules.py:5-66; the torchvision augmentations (use_augs) with explicit parameters through ``apply_augs``."""

    def __init__(self, cut_size, num_cutouts, cutout_size_power=1.0, use_augs=False):
        super().__init__()
        self.cut_size, self.cutn, self.cut_pow = cut_size, num_cutouts, cutout_size_power
        self.cached_coords = None
        self.use_augs = use_augs

    def _generate_coords(self, side_x, side_y, cutn):  # modules.py:38-48; CPU default generator, 3 draws/cutout
        max_size = min(side_y, side_x)
        min_size = min(side_y, side_x, self.cut_size)
        coords = []
        for _ in range(cutn):
            size = int(th.rand([]) ** self.cut_pow * (max_size - min_size) + min_size)
            ox = th.randint(0, side_x - size + 1, ()).item()
            oy = th.randint(0, side_y - size + 1, ()).item()
            coords.append((ox, oy, size))
        return coords

    def cache_coordinates(self, side_x, side_y):  # modules.py:26-36
        self.cached_coords = self._generate_coords(side_x, side_y, self.cutn)

    resize = "pool"  # "lanczos3": the ResizeRight mode named by north_star (oracle/resize_right.py) instead of the reference's pooling

    def forward(self, x, use_cache=False, num_cutouts_override=None, coords=None, aug_params=None, aug_noise=None):
        """aug_params [cutn, 20] (+ aug_noise [cutn, 4, B, 3, Smax, Smax]): the use_augs pipeline with explicit randomness"""
        cutn = num_cutouts_override if num_cutouts_override is not None else self.cutn
        side_x, side_y = x.shape[2:4]  # sic: (H, W) named (x, y), modules.py:52 (quirk B3)
        if coords is None:
            if use_cache and self.cached_coords is not None:
                coords = self.cached_coords[:cutn]
            else:
                coords = self._generate_coords(side_x, side_y, cutn)
        if self.resize == "lanczos3":
            from .resize_right import resize_lanczos3
            return th.cat([resize_lanczos3(x[:, :, oy:oy + s, ox:ox + s], (self.cut_size, self.cut_size)) for ox, oy, s in coords])
        if aug_params is not None:
            outs = []
            for k, (ox, oy, s) in enumerate(coords):
                cut = x[:, :, oy:oy + s, ox:ox + s]
                nz = None if aug_noise is None else aug_noise[k][..., :cut.shape[-2], :cut.shape[-1]]
                outs.append(F.adaptive_avg_pool2d(apply_augs(cut, aug_params[k], nz), self.cut_size))
            return th.cat(outs)
        outs = [F.adaptive_avg_pool2d(x[:, :, oy:oy + s, ox:ox + s], self.cut_size) for ox, oy, s in coords]
        return th.cat(outs)


def guidance_loss(x, pred_xstart, fac, coords, clip_model, target_embeds, weights, *, cut_size,
                  clip_guidance_scale=1000.0, tv_scale=150.0, range_scale=50.0, sat_scale=0.0,
                  lpips_model=None, init_tensor=None, init_scale=0.0, cutout_resize="pool", aug_params=None, aug_noise=None):
    """The differentiable body of cond_fn (cgd/cgd.py:177-226) with explicit cutout coordinates; the LPIPS term
    (cgd/cgd.py:220-224) when ``lpips_model`` (oracle/lpips.py) and ``init_tensor`` are given.
    Returns (total_loss, dict of per-term scalars)."""
    n = x.shape[0]
    cutn = len(coords)
    x_in = pred_xstart * fac + x * (1 - fac)
    mk = MakeCutouts(cut_size, cutn)
    mk.resize = cutout_resize
    clip_in = clip_normalize(mk(x_in.add(1).div(2), coords=coords, aug_params=aug_params, aug_noise=aug_noise))
    embeds = clip_model.encode_image(clip_in).float().view([cutn, n, -1])
    dists = spherical_dist_loss(embeds.unsqueeze(0), target_embeds.unsqueeze(0)).view([cutn, n, -1])
    clip_l = dists.mul(weights).sum(2).mean(0).sum() * clip_guidance_scale
    range_l = range_loss(pred_xstart).sum() * range_scale
    tv_l = tv_loss(x_in).sum() * tv_scale
    loss = clip_l + tv_l + range_l
    terms = {"clip": clip_l, "range": range_l, "tv": tv_l}
    if sat_scale != 0:
        sat_l = th.abs(x_in - x_in.clamp(min=-1, max=1)).mean() * sat_scale
        loss = loss + sat_l
        terms["sat"] = sat_l
    if lpips_model is not None and init_tensor is not None and init_scale != 0:  # cgd/cgd.py:220-224
        init_l = lpips_model(x_in, init_tensor).sum() * init_scale
        loss = loss + init_l
        terms["init"] = init_l
    return loss, terms


def magnitude_clamp(g, max_rms=0.05):  # cgd/cgd.py:229-232
    mag = g.square().mean().sqrt()
    return g * mag.clamp(max=max_rms) / mag


class OracleCondFn:
    """Stateful stand-in for the cond_fn closure + its ``current_timestep`` bookkeeping
    (cgd/cgd.py:149-239,265-267).  ``coords_fn(cutn)`` supplies cutout windows (defaults to fresh
    CPU-RNG draws exactly like the reference)."""

    def __init__(self, diffusion, clip_model, target_embeds, weights, *, cut_size, num_cutouts,
                 cutout_power=1.0, clip_guidance_scale=1000.0, tv_scale=150.0, range_scale=50.0,
                 sat_scale=0.0, use_magnitude=False, lpips_model=None, init_tensor=None, init_scale=0.0, cutout_resize="pool"):
        self.diffusion, self.clip_model = diffusion, clip_model
        self.target_embeds, self.weights = target_embeds, weights
        self.mk = MakeCutouts(cut_size, num_cutouts, cutout_power)
        self.kw = dict(cut_size=cut_size, clip_guidance_scale=clip_guidance_scale, tv_scale=tv_scale,
                       range_scale=range_scale, sat_scale=sat_scale, lpips_model=lpips_model, init_tensor=init_tensor,
                       init_scale=init_scale, cutout_resize=cutout_resize)
        self.use_magnitude = use_magnitude
        self.current_timestep = diffusion.num_timesteps - 1
        self.last_coords = None
        self.last_terms = None
        self.log_items = True

    def step_done(self):
        self.current_timestep -= 1

    def __call__(self, x, t, out, y=None, coords=None, aug_params=None, aug_noise=None):
        fac = float(self.diffusion.sqrt_one_minus_alphas_cumprod[self.current_timestep])
        if coords is None:
            coords = self.mk._generate_coords(x.shape[2], x.shape[3], self.mk.cutn)
        self.last_coords = coords
        loss, terms = guidance_loss(x, out["pred_xstart"], fac, coords, self.clip_model, self.target_embeds,
                                    self.weights, aug_params=aug_params, aug_noise=aug_noise, **self.kw)
        # the reference logs the loss terms with .item() every step (cgd/cgd.py:234-236: three host syncs); log_items=False keeps
        # them on the device (bench.py times the PyTorch-CUDA arm both ways)
        self.last_terms = {k: float(v.detach()) for k, v in terms.items()} if self.log_items else {k: v.detach() for k, v in terms.items()}
        g = -th.autograd.grad(loss, x)[0]
        if self.use_magnitude:
            g = magnitude_clamp(g)
        return g
