"""The guided sampling step engine: UNet fwd -> cond_fn (cutouts, CLIP fwd/bwd, losses) -> UNet dgrad -> update.

Replaces the L1 <-> L3 <-> L2 ping-pong the reference executes once per timestep:
``p_mean_variance`` + ``cond_fn`` (cgd/cgd.py:151-239) + ``-autograd.grad`` (cgd/cgd.py:228) + the ancestral / DDIM
update.  One op plan holds both networks and every guidance kernel in a single arena; a whole step is ~1.3k kernel
launches replayed from one CUDA graph, fed per step by one small pinned H2D copy (scalars, cutout windows, t, y).

Surfaces kept: ``cond_fn(x, t, out, y=None)``, ``MakeCutouts``, ``model(x, timesteps, y)``.
"""
from __future__ import annotations

import numpy as np
import torch as th

from . import _lib
from ._lib import SC
from .plan import Plan
from .unet import IN_PAD, UNetB200, UNetConfig
from .rn import RNB200, RNConfig
from .vit import ViTB200, ViTConfig

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # cgd/clip_util.py:45
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class MakeCutouts(th.nn.Module):
    """cgd/modules.py:5-66 surface: same constructor, ``cached_coords``, ``cache_coordinates`` and the same CPU-generator
    draw order (3 draws per cutout: size, offsetx, offsety).  ``forward`` runs the batched CUDA cutout kernel and
    returns ``[cutn*B, 3, cut_size, cut_size]`` fp32 (un-normalised pooled cutouts, like the reference)."""

    def __init__(self, cut_size: int, num_cutouts: int, cutout_size_power: float = 1.0, use_augs: bool = False):
        super().__init__()
        self.cut_size, self.cutn, self.cut_pow = cut_size, num_cutouts, cutout_size_power
        self.cached_coords = None
        # use_augs (cgd/modules.py:12-24): flip / affine / perspective / grayscale / noise run inside the cutout kernels
        # (csrc/augs.cu) from parameters drawn in torchvision's order (augs.py); `augs` stays an attribute like the reference's
        self.use_augs = bool(use_augs)
        self.augs = th.nn.Identity()

    def _generate_coords(self, side_x: int, side_y: int, cutn: int):
        max_size = min(side_y, side_x)
        min_size = min(side_y, side_x, self.cut_size)
        coords = []
        for _ in range(cutn):
            size = int(th.rand([]) ** self.cut_pow * (max_size - min_size) + min_size)
            offsetx = th.randint(0, side_x - size + 1, ()).item()
            offsety = th.randint(0, side_y - size + 1, ()).item()
            coords.append((offsetx, offsety, size))
        return coords

    def cache_coordinates(self, side_x: int, side_y: int):
        self.cached_coords = self._generate_coords(side_x, side_y, self.cutn)

    def coords_for(self, side_x, side_y, use_cache=False, num_cutouts_override=None):
        cutn = num_cutouts_override if num_cutouts_override is not None else self.cutn
        if use_cache and self.cached_coords is not None:
            if len(self.cached_coords) < cutn:  # the reference fails one line later, at .view([current_cutn, n, -1]) (cgd/cgd.py:194-195)
                raise RuntimeError(f"{cutn} cutouts requested but only {len(self.cached_coords)} cached: cached_cutouts with "
                                   "progressive_cutout needs num_cutouts >= 16 (the schedule's middle count is max(8, n // 2))")
            return self.cached_coords[:cutn]
        return self._generate_coords(side_x, side_y, cutn)

    def forward(self, input: th.Tensor, use_cache: bool = False, num_cutouts_override: int = None):
        import ctypes
        if not input.is_cuda:
            raise _lib.CgdError("MakeCutouts.forward runs on the GPU only (no CPU fallback)")
        B, C, H, W = input.shape
        assert C == 3
        coords = self.coords_for(H, W, use_cache, num_cutouts_override)  # sic: (H, W) passed as (side_x, side_y), modules.py:52
        cs = self.cut_size
        x = input.detach().float().contiguous()
        cdev = th.tensor(coords, dtype=th.int32, device=x.device)
        lib = _lib.load()
        if self.use_augs:  # the reference applies its augmentations to the raw cutout values: feed 2x - 1, mean 0 / std 1
            from . import augs
            out = th.empty(len(coords) * B, 1, 3 * cs * cs, dtype=th.float16, device=x.device)
            Smax = min(H, W)
            noise = th.zeros(len(coords), 4, B, 3, Smax, Smax, device=x.device)
            prm = augs.draw_aug_params(coords, B, H, W, noise_device=x.device, noise_out=noise).to(x.device)
            xin = x * 2 - 1
            op = _lib.CgdOp()
            op.code = _lib.OP["CUTOUTS_AUG_FWD"]
            for j, v in enumerate([B, H, W, len(coords), cs, cs, 3 * cs * cs, Smax]):
                op.i[j] = v
            for j, v in enumerate([0.0, 0.0, 0.0, 1.0, 1.0, 1.0]):
                op.f[j] = v
            for j, t in enumerate([xin, cdev, out, prm, noise]):
                op.p[j] = t.data_ptr()
            _lib.check(lib.cgd_run_op(ctypes.byref(op), ctypes.c_void_p(th.cuda.current_stream().cuda_stream)), "cutouts_aug_fwd")
            return out.view(len(coords) * B, 3, cs, cs).float()
        # mean 0.5 / std 0.5 with the kernel's (x + 1) / 2 yield the raw pooled value; flags 1 = fp32 output like the reference's
        # adaptive_avg_pool2d (the fused step writes fp16 straight into the CLIP tower's patch buffer instead)
        out32 = th.empty(len(coords) * B, 3, cs, cs, dtype=th.float32, device=x.device)
        op = _lib.CgdOp()
        op.code, op.flags = _lib.OP["CUTOUTS_FWD"], 1
        for j, v in enumerate([B, H, W, len(coords), cs, cs, 3 * cs * cs]):
            op.i[j] = v
        for j, v in enumerate([0.5, 0.5, 0.5, 0.5, 0.5, 0.5]):
            op.f[j] = v
        for j, t in enumerate([x, cdev, out32]):
            op.p[j] = t.data_ptr()
        _lib.check(lib.cgd_run_op(ctypes.byref(op), ctypes.c_void_p(th.cuda.current_stream().cuda_stream)), "cgd_cutouts_fwd")
        return out32


class EngineModel:
    """What the loops and ``p_sample`` receive as ``model``: ``model(x, timesteps, y) -> [B, 6, H, W]``."""

    def __init__(self, engine):
        self.engine = engine
        self.num_classes = engine.unet.num_classes
        self.dtype = th.float16

    def __call__(self, x, timesteps, y=None):
        return self.engine.unet(x, timesteps, y)

    def parameters(self):
        yield self.engine.plan.arena

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    def convert_to_fp16(self):
        return self

    def to(self, device=None, *a, **k):  # cgd/script_util.py:318 chains .requires_grad_(False).eval().to(device)
        if device is not None and th.device(device).type != self.engine.device.type:
            raise RuntimeError(f"the engine lives on {self.engine.device}; build a GuidedStepB200 per device")
        return self

    def load_state_dict(self, state_dict, strict=True):
        """cgd/script_util.py:317 ``model.load_state_dict(th.load(checkpoint_path))``: re-pack a checkpoint in upstream key layout
        into the kernel layouts (K-major fp16 tiles, tap-flipped dgrad copies, the fused emb_layers matrix) IN PLACE -- the engine's
        op lists, TMA descriptors and captured CUDA graphs hold addresses and stay valid.  The packing code runs once more on the
        host against a shadow plan (seconds, like construction); shapes must match the architecture the engine was built for."""
        from . import weights as W
        from .plan import Plan
        eng = self.engine
        want = W.unet_param_shapes(eng.unet.cfg)
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in want]
        bad = [k for k in want if k in state_dict and tuple(state_dict[k].shape) not in (tuple(want[k]), tuple(want[k]) + (1,))]
        if bad:
            raise RuntimeError("size mismatch for " + ", ".join(f"{k}: {tuple(state_dict[k].shape)} vs {tuple(want[k])}" for k in bad[:4]))
        if missing or (strict and unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:4]}{'...' if len(missing) > 4 else ''}, "
                               f"unexpected {unexpected[:4]}{'...' if len(unexpected) > 4 else ''}")
        shadow = Plan(conv_impl=eng.plan.conv_impl)
        for flag in ("grid_gn", "fused_gn", "tc_attention", "cluster_splitk", "gn_epi_stats"):
            setattr(shadow, flag, getattr(eng.plan, flag))
        UNetB200(eng.unet.cfg, state_dict, batch=eng.B, height=eng.H, width=eng.W, device="cpu", seed_scale=eng.unet.seed_scale, plan=shadow,
                 build_backward=eng.vit is not None)
        eng.plan.reload_consts(shadow, first=0)  # the UNet is the first network of the engine's plan
        if eng.device.type == "cuda":
            th.cuda.synchronize(eng.device)
        import collections
        return collections.namedtuple("IncompatibleKeys", "missing_keys unexpected_keys")(missing, unexpected)


class GuidedStepB200:
    """Engine for one (local batch, H, W, cutn, CLIP tower) configuration on one GPU."""

    def __init__(self, unet_cfg: UNetConfig, unet_sd: dict, vit_cfg: ViTConfig = None, vit_sd: dict = None, *, batch: int,
                 height: int = None, width: int = None, num_cutouts: int = 16, max_prompts: int = 1, clip_guidance_scale=1000.0,
                 tv_scale=150.0, range_scale=50.0, sat_scale=0.0, use_magnitude=False, device="cuda", seed_scale=0.0,
                 vit_grad_scale=1.0, conv_impl=0, rank: int = 0, world_size: int = 1, use_graph: bool = True, vit_streams: int = 1,
                 cutn_variants: tuple = (), lpips_sd: dict = None, init_scale: float = 0.0, cutout_resize: str = "pool",
                 use_augs: bool = False):
        self.device = th.device(device)
        self.B = batch
        self.rank, self.world = rank, world_size
        self.global_batch = batch * world_size
        self.H = height or unet_cfg.image_size
        self.W = width or unet_cfg.image_size
        self.cutn = num_cutouts if vit_cfg is not None else 0
        self.P = max_prompts
        self.scales = dict(cgs=float(clip_guidance_scale), tv=float(tv_scale), rng=float(range_scale), sat=float(sat_scale))
        self.use_magnitude = bool(use_magnitude)
        self.mag_sync = False
        self.seed_scale, self.vit_grad_scale = float(seed_scale), float(vit_grad_scale)
        self.use_graph = use_graph and self.device.type == "cuda"
        self.plan = p = Plan(conv_impl=conv_impl)
        B, H, W, HW = self.B, self.H, self.W, self.H * self.W
        self.unet = UNetB200(unet_cfg, unet_sd, batch=B, height=H, width=W, device=device, seed_scale=seed_scale or 1.0, plan=p,
                             build_backward=vit_cfg is not None)
        n3 = B * 3 * HW
        self.sc = p.new(SC["COUNT"], "f", "scalars")
        self.noise = p.new(n3, "f", "noise")
        self.x0 = p.new(n3, "f", "pred_xstart")
        self.mean = p.new(n3, "f", "mean")
        self.var = p.new(n3, "f", "variance")
        self.logvar = p.new(n3, "f", "log_variance")
        self.x_inb = p.new(n3, "f", "x_in")
        self.g = p.new(n3, "f", "g")
        self.sample = p.new(n3, "f", "sample")
        self.loss = p.new(4 * B, "f", "losses")  # [clip | tv | range | sat] per image
        p.mark("pmv")
        p.emit("PMV_BLEND", i=[B, HW, 3 * B], p=[(self.unet.x_in, 0), (self.unet.out, 0), (self.sc, 0), (self.x0, 0), (self.mean, 0),
                                                 (self.var, 0), (self.logvar, 0), (self.x_inb, 0), (self.loss, B)], tag="p_mean_variance+blend")
        p.mark("cond")
        self.vit = None
        self.lpips = None
        self.use_augs = False
        if vit_cfg is not None:
            cutn, cs, ps, kp, D = self.cutn, vit_cfg.input_resolution, vit_cfg.patch_size, vit_cfg.kpad, vit_cfg.output_dim
            self.vit_cfg = vit_cfg
            self.coords = p.new(cutn * 3, "i32", "cutout_coords")
            self.targets = p.new(self.P * D, "f", "target_embeds")
            self.weights = p.new(self.P, "f", "prompt_weights")
            self.g_clip = p.new(n3, "f", "g_clip")
            self.dx_direct = p.new(n3, "f", "dx_direct")
            self.fg_ws = p.new(128, "f", "final_grad_ws")
            # CLIP visual tower: ViT (vit.py) or ModifiedResNet (rn.py); both expose patches / embeds / d_embeds / d_patches
            Tower = RNB200 if isinstance(vit_cfg, RNConfig) else ViTB200
            self.vit = Tower(vit_cfg, vit_sd, n_images=cutn * B, device=device, plan=p, parts=vit_streams)
            # further cutout counts of the same engine (progressive_cutout, cgd/cgd.py:167-175): own activations and op ranges
            # ("...@c" marks), shared packed weights; the default count keeps the un-suffixed marks
            self.vits = {cutn: self.vit}
            # cutout resampling: "pool" = adaptive_avg_pool2d like the reference's MakeCutouts (cgd/modules.py:63); "lanczos3" =
            # the vendored ResizeRight resampler north_star names (cgd/ResizeRight/resize_right.py, host tables in resize_right.py)
            if cutout_resize not in ("pool", "lanczos3"):
                raise ValueError("cutout_resize must be 'pool' or 'lanczos3'")
            self.cutout_resize = cutout_resize
            self.use_augs = bool(use_augs)
            if self.use_augs:
                if cutout_resize != "pool":
                    raise NotImplementedError("use_augs runs with the reference's adaptive_avg_pool2d cutouts (cutout_resize='pool')")
                from .augs import AUG_NP
                self.aug_smax = min(H, W)
                self.aug_prm = p.new(cutn * AUG_NP, "f", "aug_params")
                self.aug_noise = p.new(cutn * 4 * B * 3 * self.aug_smax * self.aug_smax, "f", "aug_noise")
            if cutout_resize == "lanczos3":
                if H != W or cutn_variants:
                    raise NotImplementedError("ResizeRight cutouts: square images and a single cutout count only")
                from .resize_right import T_MAX
                # the device tables hold T_MAX = 16 taps per output: lanczos3 with antialiasing needs ceil(6 * S / cs) of them, and the
                # largest window the reference draws is min(H, W) (cgd/modules.py:40) -- checked here, once, not in the middle of a run
                if -(-6 * min(H, W) // cs) > T_MAX:
                    raise ValueError(f"cutout_resize='lanczos3': windows up to {min(H, W)} px down to {cs} px need {-(-6 * min(H, W) // cs)} taps, "
                                     f"the device tables hold {T_MAX} (images up to {T_MAX * cs // 6} px)")
                self.rr_left = p.new(cutn * cs, "i32", "rr_left")
                self.rr_w = p.new(cutn * cs * T_MAX, "f", "rr_weights")
                self.rr_taps = p.new(cutn, "i32", "rr_taps")
                self.rr_inv = p.new(cutn * H * 2, "i32", "rr_inverse")
            for c in sorted(set(int(v) for v in cutn_variants) - {cutn}):
                if not 0 < c < cutn:
                    raise ValueError(f"cutn_variants must lie in (0, {cutn}), got {c}")
                self.vits[c] = Tower(vit_cfg, vit_sd, n_images=c * B, device=device, plan=p, parts=vit_streams, suffix=f"@{c}", share=self.vit)
            for c in sorted(self.vits, key=lambda v: v == cutn):  # the default count last: its ranges end at "guide"
                sfx, vt = ("" if c == cutn else f"@{c}"), self.vits[c]
                p.mark("cut_fwd" + sfx)
                if cutout_resize == "lanczos3":
                    p.emit("CUTOUTS_RR_FWD", i=[B, H, W, c, cs, ps, kp], f=[*CLIP_MEAN, *CLIP_STD],
                           p=[(self.x_inb, 0), (self.coords, 0), (vt.patches, 0), (self.rr_left, 0), (self.rr_w, 0), (self.rr_taps, 0)],
                           tag="make_cutouts(resize_right lanczos3)+normalize")
                elif self.use_augs:
                    p.emit("CUTOUTS_AUG_FWD", i=[B, H, W, c, cs, ps, kp, self.aug_smax], f=[*CLIP_MEAN, *CLIP_STD],
                           p=[(self.x_inb, 0), (self.coords, 0), (vt.patches, 0), (self.aug_prm, 0), (self.aug_noise, 0)],
                           tag="make_cutouts(use_augs)+normalize")
                else:
                    p.emit("CUTOUTS_FWD", i=[B, H, W, c, cs, ps, kp], f=[*CLIP_MEAN, *CLIP_STD], p=[(self.x_inb, 0), (self.coords, 0), (vt.patches, 0)],
                           tag="make_cutouts+normalize")
                p.mark("sph" + sfx)
                p.emit("SPHERICAL", i=[c, B, self.P, D], f=[self.scales["cgs"], self.vit_grad_scale],
                       p=[(vt.embeds, 0), (self.targets, 0), (self.weights, 0), (vt.d_embeds, 0), (self.loss, 0)], tag="spherical_dist_loss")
                p.mark("cut_bwd" + sfx)
                if cutout_resize == "lanczos3":
                    p.emit("CUTOUTS_RR_BWD", i=[B, H, W, c, cs, ps, kp, H], f=[0, 0, 0, *CLIP_STD, 1.0 / self.vit_grad_scale],
                           p=[(vt.d_patches, 0), (self.coords, 0), (self.g_clip, 0), (self.rr_left, 0), (self.rr_w, 0), (self.rr_inv, 0)],
                           tag="d_make_cutouts(resize_right)")
                elif self.use_augs:
                    p.emit("FILL", i=[n3], f=[0.0], p=[(self.g_clip, 0)], tag="zero g_clip")
                    p.emit("CUTOUTS_AUG_BWD", i=[B, H, W, c, cs, ps, kp], f=[0, 0, 0, *CLIP_STD, 1.0 / self.vit_grad_scale],
                           p=[(vt.d_patches, 0), (self.coords, 0), (self.g_clip, 0), (self.aug_prm, 0)], tag="d_make_cutouts(use_augs)")
                else:
                    p.emit("CUTOUTS_BWD", i=[B, H, W, c, cs, ps, kp], f=[0, 0, 0, *CLIP_STD, 1.0 / self.vit_grad_scale],
                           p=[(vt.d_patches, 0), (self.coords, 0), (self.g_clip, 0)], tag="d_make_cutouts")
                p.mark("cut_end" + sfx)
            if lpips_sd is not None and init_scale != 0:
                from .lpips import LpipsB200
                # d (init_scale * lpips_vgg(x_in, init).sum()) / d x_in is added to the x_in gradient the cutout backward left in g_clip
                self.lpips = LpipsB200(lpips_sd, B, H, W, p, self.x_inb, self.g_clip, init_scale)
            p.mark("guide")
            # seed_scale > 0: static loss scale; seed_scale = 0 (default): per-image power-of-two scale chosen on the device every step
            dyn = self.seed_scale == 0.0
            if dyn:
                self.seed_f32 = p.new(n3, "f", "seed_f32")
                self.seed_dyn = p.new(2 * B, "f", "seed_dyn")
            p.emit("GUIDE_GRAD", flags=1 if dyn else 0, i=[B, H, W, IN_PAD, self.global_batch],
                   f=[self.scales["tv"], self.scales["rng"], self.scales["sat"], self.seed_scale],
                   p=[(self.x_inb, 0), (self.x0, 0), (self.g_clip, 0), (self.sc, 0), (self.unet.seed, 0), (self.dx_direct, 0), (self.loss, B)]
                   + ([(self.seed_f32, 0), (self.seed_dyn, 0)] if dyn else []), tag="tv+range+sat")
            if dyn:
                p.emit("SEED_QUANT", i=[B, HW, IN_PAD], p=[(self.seed_f32, 0), (self.seed_dyn, 0), (self.unet.seed, 0)], tag="seed scale")
            p.mark("final")
            # use_magnitude clamps by the RMS of the WHOLE batch (cgd/cgd.py:229-232).  With the batch sharded over ranks FINAL_GRAD
            # only writes its partial sums, the host all-reduces the 128 floats (the step's one data-path collective, 512 bytes)
            # and MAG_CLAMP applies the clamp; a single rank keeps the two-launch FINAL_GRAD.
            self.mag_sync = self.use_magnitude and self.world > 1
            p.emit("FINAL_GRAD", flags=(1 if self.use_magnitude else 0) | (2 if dyn else 0) | (4 if self.mag_sync else 0), i=[B, HW],
                   f=[1.0 / self.seed_scale if not dyn else 0.0, 0.05],
                   p=[(self.dx_direct, 0), (self.unet.dx, 0), (self.g, 0), (self.fg_ws, 0)] + ([(self.seed_dyn, 0)] if dyn else []), tag="-grad")
            if self.mag_sync:
                p.mark("mag")
                p.emit("MAG_CLAMP", i=[n3, self.global_batch * 3 * HW], f=[0.05], p=[(self.g, 0), (self.fg_ws, 0)], tag="rms clamp")
        n = n3
        p.mark("upd_anc_g")
        p.emit("SAMPLE_ANCESTRAL", i=[n], p=[(self.mean, 0), (self.var, 0), (self.logvar, 0), (self.g, 0), (self.noise, 0), (self.sc, 0), (self.sample, 0)])
        p.mark("upd_anc")
        p.emit("SAMPLE_ANCESTRAL", i=[n], p=[(self.mean, 0), (self.var, 0), (self.logvar, 0), None, (self.noise, 0), (self.sc, 0), (self.sample, 0)])
        p.mark("upd_ddim_g")
        p.emit("SAMPLE_DDIM", i=[n], p=[(self.unet.x_in, 0), (self.x0, 0), (self.g, 0), (self.noise, 0), (self.sc, 0), (self.sample, 0)])
        p.mark("upd_ddim")
        p.emit("SAMPLE_DDIM", i=[n], p=[(self.unet.x_in, 0), (self.x0, 0), None, (self.noise, 0), (self.sc, 0), (self.sample, 0)])
        p.mark("engine_end")
        p.finalize(device)
        self.model = EngineModel(self)
        self.shape = (B, 3, H, W)
        # pinned staging for the per-step host->device refresh
        self._n_stage = SC["COUNT"] * 4 + self.cutn * 3 * 4 + B * 4 + B * 8 + self.cutn * 20 * 4
        # The host runs ahead of the GPU (a step is ~1 ms of host work and ~10 ms of device work, and nothing synchronises between
        # saved frames), so the async copies of step k may still be queued when the host stages step k+1: the pinned buffers are a
        # ring of STAGE_SLOTS slots, each guarded by the event recorded after its copies were enqueued.
        pin = self.device.type == "cuda"
        self._stage_ring = [th.zeros(self._n_stage, dtype=th.uint8, pin_memory=pin) for _ in range(self.STAGE_SLOTS)]
        self._stage_events = [None] * self.STAGE_SLOTS
        self._stage_i = 0
        self._rr_ring = None
        self._graphs = {}
        self._side_streams = [th.cuda.Stream(device=self.device) for _ in range((self.vit.parts if self.vit is not None else 1) - 1)] \
            if self.device.type == "cuda" else []
        self._last_out = None
        self._fwd_valid = False
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ------------------------------------------------------------------ views
    def v(self, buf, shape=None):
        return self.plan.view(buf, shape)

    def img(self, buf):
        return self.plan.view(buf, self.shape)

    def set_init_image(self, init: th.Tensor, runner=None):
        """init image in [-1, 1] for the LPIPS init loss (cgd/cgd.py:116-120, 220-224); engines built with lpips_sd only"""
        if self.lpips is None:
            raise RuntimeError("engine built without the LPIPS loss (pass lpips_sd= and init_scale != 0)")
        self.lpips.set_init_image(init, runner)

    def set_targets(self, target_embeds: th.Tensor, weights: th.Tensor):
        """target_embeds [P, D] (un-normalised is fine: the loss normalises, cgd/losses.py:12-13); weights already divided by
        their |sum| like cgd/cgd.py:102-105."""
        P = target_embeds.shape[0]
        if P != self.P:
            raise ValueError(f"engine was built for {self.P} prompt(s), got {P}")
        if self.B > 1 and P > 1:
            raise RuntimeError("the reference's spherical-loss broadcast is only defined for batch 1 or a single prompt (quirk B1)")
        self.v(self.targets, (P, self.vit_cfg.output_dim)).copy_(target_embeds.detach().float())
        self.v(self.weights, (P,)).copy_(weights.detach().float())

    # ------------------------------------------------------------------ RNG (torch generators: seed parity)
    def local_rows(self, t: th.Tensor) -> th.Tensor:
        return t[self.rank * self.B:(self.rank + 1) * self.B]

    def draw_initial_noise(self, shape=None) -> th.Tensor:
        full = th.randn(self.global_batch, 3, self.H, self.W, device=self.device)
        return self.local_rows(full).contiguous()

    def draw_noise(self) -> th.Tensor:
        """th.randn_like(x) of the reference; with several ranks every rank draws the full batch and keeps its rows."""
        if self.world == 1:
            return self.img(self.noise).normal_()
        full = th.randn(self.global_batch, 3, self.H, self.W, device=self.device)
        self.img(self.noise).copy_(self.local_rows(full))
        return self.img(self.noise)

    def gather_final(self, t: th.Tensor) -> th.Tensor:
        """The run's single collective: every rank's rows of ``t`` ([B_local, ...], e.g. the final sample or
        pred_xstart) -> [B_global, ...] in rank order on every rank.  One ``all_gather_into_tensor`` (NCCL over NVLink; gloo in the
        CPU tests) into a buffer allocated once per shape -- no list of outputs, no copies, no per-call channel set-up after the
        first.  A single rank returns its tensor unchanged."""
        if self.world == 1:
            return t
        import torch.distributed as dist
        t = t.contiguous()
        key = (tuple(t.shape), t.dtype)
        bufs = self.__dict__.setdefault("_gather_bufs", {})
        out = bufs.get(key)
        if out is None:
            out = bufs[key] = th.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out

    def draw_classes(self) -> th.Tensor:
        full = th.randint(0, self.unet.num_classes, (self.global_batch,), device=self.device)
        return self.local_rows(full)

    # ------------------------------------------------------------------ segment-wise execution (API-parity path)
    def _push_scalars(self, sc: np.ndarray):
        self.v(self.sc).copy_(th.from_numpy(sc), non_blocking=True)

    def unet_forward(self, diffusion, x, t_index: int, y=None, fac_index=None) -> dict:
        if fac_index is None:
            fac_index = t_index
        self._cur_t = t_index
        self._push_scalars(diffusion.scalar_table(t_index, fac_index))
        ts = th.full((self.B,), diffusion.model_timestep(t_index), dtype=th.float32)
        self.unet.set_inputs(x, ts.to(self.device, non_blocking=True), y)
        self.plan.run_range("unet_emb", "unet_bwd")
        self.plan.run_range("pmv", "cond")
        self._fwd_valid = True
        out = {"mean": self.img(self.mean), "variance": self.img(self.var), "log_variance": self.img(self.logvar),
               "pred_xstart": self.img(self.x0), "engine": self}
        self._last_out = out
        return out

    def cond_grad(self, diffusion, coords, fac_index: int) -> th.Tensor:
        """-d(loss)/dx for the last unet_forward (what cond_fn returns)."""
        if not self._fwd_valid:
            raise RuntimeError("cond_fn called without a preceding p_mean_variance on this engine")
        if self.vit is None:
            raise RuntimeError("engine built without a CLIP tower")
        if fac_index != self._cur_t:  # quirk B2: fac follows the closure's current_timestep, not t
            sc = diffusion.scalar_table(self._cur_t, fac_index)
            self._push_scalars(sc)
            self.plan.run_range("pmv", "cond")
        cutn = len(coords)
        sfx = self._sfx(cutn)
        self.v(self.coords, (self.cutn, 3))[:cutn].copy_(th.tensor(coords, dtype=th.int32), non_blocking=True)
        self._stage_resize_tables(coords)
        if self.use_augs:
            self._stage_augs(coords)
        self.plan.run_range("cut_fwd" + sfx, "sph" + sfx)
        self._run_vit("fwd", None, cutn)
        self.plan.run_range("sph" + sfx, "cut_bwd" + sfx)
        self._run_vit("bwd", None, cutn)
        for a, b in (("cut_bwd" + sfx, "cut_end" + sfx),) + ((("lpips", "lpips_end"),) if self.lpips is not None else ()) + (
                ("guide", "final"), ("unet_bwd", "unet_end")):
            self.plan.run_range(a, b)
        self._run_final(self.plan.run_range)
        return self.img(self.g)

    def _run_final(self, pr):
        """-grad (+ RMS clamp).  Sharded batch with use_magnitude: partial sums -> all-reduce over ranks -> clamp."""
        if not self.mag_sync:
            pr("final", "upd_anc_g")
            return
        import torch.distributed as dist
        pr("final", "mag")
        dist.all_reduce(self.v(self.fg_ws))
        pr("mag", "upd_anc_g")

    def update(self, diffusion, mode, t_index, g, noise, eta=0.0) -> th.Tensor:
        if mode == "ddim" and eta != 0.0:
            self.v(self.sc)[SC["ETA"]] = float(eta)
        if noise.data_ptr() != self.v(self.noise).data_ptr():
            self.img(self.noise).copy_(noise)
        if g is not None and g.data_ptr() != self.v(self.g).data_ptr():
            self.img(self.g).copy_(g)
        name = ("upd_anc" if mode == "ancestral" else "upd_ddim") + ("_g" if g is not None else "")
        order = ["upd_anc_g", "upd_anc", "upd_ddim_g", "upd_ddim", "engine_end"]
        self.plan.run_range(name, order[order.index(name) + 1])
        self._fwd_valid = False
        return self.img(self.sample).clone()

    # ------------------------------------------------------------------ fused step (one CUDA graph)
    def can_fuse(self, cond_fn, clip_denoised, denoised_fn) -> bool:
        return (isinstance(cond_fn, CondFnB200) and cond_fn.engine is self and not clip_denoised and denoised_fn is None
                and cond_fn.fusable())

    def _sfx(self, cutn):
        if cutn is None or cutn == self.cutn:
            return ""
        if cutn not in self.vits:
            raise ValueError(f"engine built for cutout counts {sorted(self.vits)}, got {cutn}")
        return f"@{cutn}"

    def _run_all(self, mode, runner=None, cutn=None, part=None):
        pr = runner or self.plan.run_range
        sfx = self._sfx(cutn)
        pr("unet_emb", "unet_bwd")
        pr("pmv", "cond")
        pr("cut_fwd" + sfx, "sph" + sfx)
        self._run_vit("fwd", pr, cutn)
        pr("sph" + sfx, "cut_bwd" + sfx)
        self._run_vit("bwd", pr, cutn)
        pr("cut_bwd" + sfx, "cut_end" + sfx)
        if self.lpips is not None:
            pr("lpips", "lpips_end")
        pr("guide", "final")
        pr("unet_bwd", "unet_end")
        if part == "A":  # everything up to the partial sums of the sharded RMS clamp (see replay)
            pr("final", "mag")
            return
        self._run_final(pr)
        self._run_update(mode, pr)

    def _run_update(self, mode, pr):
        if mode == "ancestral":
            pr("upd_anc_g", "upd_anc")
        else:
            pr("upd_ddim_g", "upd_ddim")

    def _run_vit(self, which, pr=None, cutn=None):
        """CLIP forward / backward: one op range, or one per batch slice on parallel streams (fork / join around the current
        stream; inside CUDA-graph capture this becomes parallel branches of the graph)."""
        pr = pr or self.plan.run_range
        ranges = self.vits[self.cutn if cutn is None else cutn].part_ranges(which)
        if len(ranges) == 1 or not self._side_streams or pr != self.plan.run_range:
            for a, b in ranges:
                pr(a, b)
            return
        main = th.cuda.current_stream()
        for s, (a, b) in zip(self._side_streams, ranges[1:]):
            s.wait_stream(main)
            self.plan.run_range(a, b, stream=s.cuda_stream)
        self.plan.run_range(*ranges[0])
        for s in self._side_streams:
            main.wait_stream(s)

    def launches_per_step(self, mode="ddim") -> int:
        m = self.plan.marks
        segs = [("unet_emb", "unet_bwd"), ("pmv", "cond"), ("cut_fwd", "sph"), ("vit_fwd", "vit_bwd"), ("sph", "cut_bwd"), ("vit_bwd", "vit_end"),
                ("cut_bwd", "cut_end")] + ([("lpips", "lpips_end")] if self.lpips is not None else []) + [
                ("guide", "final"), ("unet_bwd", "unet_end"), ("final", "upd_anc_g"),
                ("upd_anc_g", "upd_anc") if mode == "ancestral" else ("upd_ddim_g", "upd_ddim")]  # "mag" lies inside (final, upd_anc_g)
        return sum(self.plan.num_launches(m[a], m[b] - m[a]) for a, b in segs)

    STAGE_SLOTS = 4

    def _stage_acquire(self):
        """next slot of the pinned ring; blocks (host only) until the copies last issued from it have executed"""
        k = self._stage_i % self.STAGE_SLOTS
        self._stage_i += 1
        ev = self._stage_events[k]
        if ev is not None:
            ev.synchronize()
        return k

    def _stage_release(self, k):
        if self.device.type == "cuda":
            ev = self._stage_events[k] or th.cuda.Event()
            ev.record(th.cuda.current_stream(self.device))
            self._stage_events[k] = ev

    def stage_step(self, sc: np.ndarray, coords, t_model: float, y=None):
        """One pinned staging slot -> small async H2D copies (classes, scalars, timestep, cutout windows)."""
        k = self._stage_acquire()
        st = self._stage_ring[k]
        o = self.B * 8  # [0, 8B): int64 classes (kept first for alignment)
        self.h2d_bytes = 0
        if y is not None and self.unet.cfg.class_cond:
            if y.device.type == "cpu":
                st[0:o].view(th.int64).copy_(y)
                self.v(self.unet.y_in).view(th.uint8).copy_(st[0:o], non_blocking=True)
                self.h2d_bytes += o
            else:
                self.v(self.unet.y_in).copy_(y)
        n = SC["COUNT"] * 4
        st[o:o + n].view(th.float32).copy_(th.from_numpy(sc))
        self.v(self.sc).view(th.uint8).copy_(st[o:o + n], non_blocking=True)
        o += n
        n = self.B * 4
        st[o:o + n].view(th.float32).fill_(t_model)
        self.v(self.unet.t_in).view(th.uint8).copy_(st[o:o + n], non_blocking=True)
        o += n
        self.h2d_bytes += SC["COUNT"] * 4 + self.B * 4
        if self.cutn:
            for ox, oy, size in coords:  # windows may be clipped at the border (quirk B3) but not empty
                if not (0 <= ox < self.W and 0 <= oy < self.H and size > 0):
                    raise RuntimeError(f"cutout window (x={ox}, y={oy}, size={size}) lies outside the {self.H}x{self.W} image: the reference's "
                                       "adaptive_avg_pool2d raises on the empty crop (non-square images, cgd/modules.py:52,61)")
            n = len(coords) * 12
        if self.cutn and len(coords):
            st[o:o + n].view(th.int32).copy_(th.tensor(coords, dtype=th.int32).view(-1))
            self.v(self.coords).view(th.uint8)[:n].copy_(st[o:o + n], non_blocking=True)
            self.h2d_bytes += n
            o += self.cutn * 12
            self._stage_resize_tables(coords, k)
            if getattr(self, "use_augs", False):
                self._stage_augs(coords, st, o)
        self._stage_release(k)

    def _stage_augs(self, coords, st=None, o=0):
        """use_augs: this step's flip / affine / perspective / grayscale parameters (CPU generator, torchvision's order) and the four
        noise fields per cutout (device generator, the reference's shapes; full batch drawn, own rows kept when sharded)"""
        from . import augs
        n = len(coords) * augs.AUG_NP * 4
        noise = self.v(self.aug_noise, (self.cutn, 4, self.B, 3, self.aug_smax, self.aug_smax))
        rows = None if self.world == 1 else (self.rank * self.B, (self.rank + 1) * self.B)
        prm = augs.draw_aug_params(coords, self.global_batch, self.H, self.W, noise_device=self.device, noise_out=noise, rows=rows)
        dst = self.v(self.aug_prm).view(th.uint8)[:n]
        if st is None:
            dst.copy_(prm.view(-1).view(th.uint8))
        else:
            st[o:o + n].view(th.float32).copy_(prm.view(-1))
            dst.copy_(st[o:o + n], non_blocking=True)
        self.h2d_bytes += n

    def _stage_resize_tables(self, coords, slot=None):
        """ResizeRight mode: per-cutout resampling tables of this step's crop sizes (cached per size on the host) -> device"""
        if self.vit is None or self.cutout_resize != "lanczos3":
            return
        from . import resize_right as rr
        cs, S_max = self.vit_cfg.input_resolution, self.H
        if self._rr_ring is None:
            pin = self.device.type == "cuda"
            self._rr_ring = [(th.zeros(self.cutn, cs, dtype=th.int32, pin_memory=pin), th.zeros(self.cutn, cs, rr.T_MAX, pin_memory=pin),
                              th.zeros(self.cutn, dtype=th.int32, pin_memory=pin), th.zeros(self.cutn, S_max, 2, dtype=th.int32, pin_memory=pin))
                             for _ in range(self.STAGE_SLOTS)]
        own = slot is None  # segment path (cond_grad): take and guard a slot of its own
        if own:
            slot = self._stage_acquire()
        pl, pw, pt, pi_ = self._rr_ring[slot]
        for k, (_, _, S) in enumerate(coords):
            left, w, T = rr.tables(int(S), cs)
            pl[k].copy_(left)
            pw[k].copy_(w)
            pt[k] = T
            pi_[k, :S].copy_(rr.inverse_ranges(left, T, int(S)))
        self.v(self.rr_left, (self.cutn, cs)).copy_(pl, non_blocking=True)
        self.v(self.rr_w, (self.cutn, cs, rr.T_MAX)).copy_(pw, non_blocking=True)
        self.v(self.rr_taps, (self.cutn,)).copy_(pt, non_blocking=True)
        self.v(self.rr_inv, (self.cutn, S_max, 2)).copy_(pi_, non_blocking=True)
        self.h2d_bytes += pl.numel() * 4 + pw.numel() * 4 + pt.numel() * 4 + pi_.numel() * 4
        if own:
            self._stage_release(slot)

    def fused_step(self, diffusion, mode, t_index, img, y, cond_fn, eta=0.0) -> dict:
        # RNG order of the reference: the ancestral sampler draws its noise BEFORE cond_fn (whose MakeCutouts draws the windows
        # from the CPU generator), DDIM AFTER it.  On a GPU the two come from different generators and the order is immaterial;
        # it is kept anyway so a CPU run (tests/test_loops_cpu.py) consumes the default generator exactly like the reference.
        fac_index = cond_fn.current_timestep
        # reduce_clip (cgd/cgd.py:157-164): on the steps its rule skips, cond_fn returns zeros -- no cutout windows are drawn, CLIP
        # and both backward passes have nothing to do, the sampler takes its unguided update.  Those steps replay a second, short
        # graph (UNet forward -> p_mean_variance -> update), which is where the option's speed-up comes from.
        guided = not cond_fn.skips_guidance()
        if mode == "ancestral":
            self.draw_noise()
        coords = cond_fn.next_coords(self.H, self.W) if guided else []
        if mode != "ancestral":
            self.draw_noise()
        sc = diffusion.scalar_table(t_index, fac_index, eta)
        self.stage_step(sc, coords, diffusion.model_timestep(t_index), y)
        xin = self.img(self.unet.x_in)
        if img.data_ptr() != xin.data_ptr():
            xin.copy_(img, non_blocking=True)
        self.replay(mode, len(coords) if (self.cutn and guided) else None, guided=guided)
        return {"sample": self.img(self.sample).clone(), "pred_xstart": self.img(self.x0).clone()}

    def _run_unguided(self, mode, pr=None):
        """a step whose guidance gradient is zero (reduce_clip's skipped steps): UNet forward, p_mean_variance, plain update"""
        pr = pr or self.plan.run_range
        pr("unet_emb", "unet_bwd")
        pr("pmv", "cond")
        if mode == "ancestral":
            pr("upd_anc", "upd_ddim_g")
        else:
            pr("upd_ddim", "engine_end")

    def replay(self, mode, cutn=None, guided=True):
        if cutn == self.cutn:
            cutn = None
        if not guided:
            if not self.use_graph:
                self._run_unguided(mode)
                return
            g = self._graphs.get((mode, "unguided"))
            if g is None:
                self._run_unguided(mode)
                th.cuda.synchronize()
                g = th.cuda.CUDAGraph()
                with th.cuda.graph(g):
                    self._run_unguided(mode)
                self._graphs[(mode, "unguided")] = g
            g.replay()
            return
        if not self.use_graph:
            self._run_all(mode, None, cutn)
            return
        if self.mag_sync:  # two graphs around the all-reduce of the RMS partial sums
            import torch.distributed as dist
            gs = self._graphs.get((mode, cutn))
            if gs is None:
                self._run_all(mode, None, cutn)
                th.cuda.synchronize()
                ga, gb = th.cuda.CUDAGraph(), th.cuda.CUDAGraph()
                with th.cuda.graph(ga):
                    self._run_all(mode, None, cutn, part="A")
                with th.cuda.graph(gb):
                    self.plan.run_range("mag", "upd_anc_g")
                    self._run_update(mode, self.plan.run_range)
                gs = self._graphs[(mode, cutn)] = (ga, gb)
            gs[0].replay()
            dist.all_reduce(self.v(self.fg_ws))
            gs[1].replay()
            return
        g = self._graphs.get((mode, cutn))
        if g is None:
            self._run_all(mode, None, cutn)  # warm-up: sets kernel attributes, touches every buffer
            th.cuda.synchronize()
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                self._run_all(mode, None, cutn)
            self._graphs[(mode, cutn)] = g
        g.replay()

    def losses(self) -> dict:
        """per-image loss terms of the last step (device->host sync; logging only, like tqdm.write at cgd/cgd.py:234-236)"""
        l = self.v(self.loss, (4, self.B)).cpu()
        d = {"clip": l[0], "tv": l[1], "range": l[2], "sat": l[3]}
        if self.lpips is not None:
            d["init"] = self.lpips.loss_value().cpu() * self.lpips.init_scale
        return d


class CondFnB200:
    """The reference's ``cond_fn(x, t, out, y=None)`` closure (cgd/cgd.py:151-239) bound to an engine, including its
    ``current_timestep`` bookkeeping (quirk B2) and the reduce_clip / progressive_cutout / cached_cutouts switches."""

    with_grad = True

    def __init__(self, engine: GuidedStepB200, diffusion, make_cutouts: MakeCutouts, *, cached_cutouts=False, reduce_clip=False,
                 progressive_cutout=False):
        self.engine, self.diffusion, self.make_cutouts = engine, diffusion, make_cutouts
        self.cached_cutouts, self.reduce_clip, self.progressive_cutout = cached_cutouts, reduce_clip, progressive_cutout
        self.current_timestep = diffusion.num_timesteps - 1  # cgd/cgd.py:265
        # the schedule follows the user's num_cutouts (make_cutouts.cutn); for num_cutouts < 16 its middle count max(8, n // 2)
        # EXCEEDS num_cutouts, so the engine is built for the largest count of the schedule (cgd.py does that)
        counts = self.progressive_counts(make_cutouts.cutn) if progressive_cutout else (make_cutouts.cutn,)
        missing = set(counts) - set(engine.vits)
        if missing and engine.cutn:
            raise ValueError(f"engine built for cutout counts {sorted(engine.vits)}; this cond_fn needs {sorted(set(counts))} "
                             "(GuidedStepB200(num_cutouts=max, cutn_variants=the others))")

    @staticmethod
    def progressive_counts(num_cutouts: int) -> tuple:
        """the three cutout counts of the reference's schedule (cgd/cgd.py:167-175)"""
        return (max(4, num_cutouts // 4), max(8, num_cutouts // 2), num_cutouts)

    def current_cutn(self) -> int:
        n = self.make_cutouts.cutn
        if not self.progressive_cutout:
            return n
        total = self.diffusion.num_timesteps
        pct = (total - self.current_timestep) / total
        lo, mid, hi = self.progressive_counts(n)
        return lo if pct < 0.3 else (mid if pct < 0.7 else hi)

    def fusable(self):
        return True

    def skips_guidance(self) -> bool:
        """reduce_clip's rule (cgd/cgd.py:157-164): below 70 % progress CLIP guidance runs on every 4th step only (the first 20 % are
        skipped altogether through skip_timesteps, cgd/cgd.py:141-144); on the other steps cond_fn returns zeros"""
        if not self.reduce_clip:
            return False
        total = self.diffusion.num_timesteps
        pct = (total - self.current_timestep) / total
        return pct < 0.7 and int((pct - 0.2) * total) % 4 != 0

    def step_done(self):  # cgd/cgd.py:267
        self.current_timestep -= 1

    def next_coords(self, H, W):
        return self.make_cutouts.coords_for(H, W, use_cache=self.cached_cutouts, num_cutouts_override=self.current_cutn())

    def __call__(self, x, t, out, y=None):
        eng = self.engine
        if out.get("engine") is not eng:
            raise RuntimeError("cond_fn: `out` must come from this engine's p_mean_variance (the UNet backward re-enters its saved "
                               "activations)")
        if self.skips_guidance():  # cgd/cgd.py:157-164
            return th.zeros_like(x)
        coords = self.next_coords(x.shape[2], x.shape[3])
        return eng.cond_grad(self.diffusion, coords, self.current_timestep)
