"""Teacher-forced single-step parity of the engine against the oracle (shared by CPU-interpreter tests, GPU tests and
__graft_entry__.smoke()).  Identical (x_t, t, y, noise, cutout windows, weights) go to both sides."""
import copy

import numpy as np
import torch as th

from clip_guided_diffusion_b200 import gaussian_diffusion as pgd
from clip_guided_diffusion_b200 import guidance as pg
from clip_guided_diffusion_b200 import unet as pu
from clip_guided_diffusion_b200 import vit as pv
from oracle import diffusion as od
from oracle import guidance as og
from oracle.clip_vit import CLIPVisualOnly, ViTConfig as OViTConfig
from oracle.unet import UNetModel, seeded_init_, tiny_config


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def cos(a, b):
    return float(th.nn.functional.cosine_similarity(a.flatten().float(), b.flatten().float(), dim=0))


def prod_unet_cfg(ocfg):
    return pu.UNetConfig(image_size=ocfg.image_size, model_channels=ocfg.model_channels, num_res_blocks=ocfg.num_res_blocks,
                         channel_mult=ocfg.channel_mult, attention_resolutions=ocfg.attention_resolutions, num_heads=ocfg.num_heads,
                         num_head_channels=ocfg.num_head_channels, class_cond=ocfg.class_cond, num_classes=ocfg.num_classes,
                         use_new_attention_order=ocfg.use_new_attention_order, rescale_timesteps=ocfg.rescale_timesteps)


def build_tiny(device, B=2, cutn=3, image=64, use_magnitude=False, sat_scale=0.0, respacing="25", conv_impl=0, use_graph=False, P=1,
               new_order=False, vit_streams=1, cutn_variants=(), run_cutn=None, init_scale=0.0, cutout_resize="pool", scales=None, hw=None, rank=0, world_size=1, tower="vit", use_augs=False, rn_width=64):
    ocfg = tiny_config(image_size=image, model_channels=64, channel_mult=(1, 2), num_res_blocks=1, attention_resolutions=(image // 2,),
                       class_cond=True, use_new_attention_order=new_order)
    ounet = seeded_init_(UNetModel(ocfg)).eval()
    if tower == "rn":  # CLIP ModifiedResNet tower (rn.py / oracle/clip_rn.py): one bottleneck per stage, 32 x 32 cutouts
        from clip_guided_diffusion_b200 import rn as prn
        from clip_guided_diffusion_b200 import weights as pw
        from oracle import clip_rn as orn
        pvit_cfg = prn.RNConfig(layers=(1, 1, 1, 1), output_dim=64, input_resolution=32, width=rn_width)
        rn_sd = pw.seeded_rn_state_dict(pvit_cfg, seed=5)
        oclip = orn.CLIPVisualRN(orn.RNConfig(layers=(1, 1, 1, 1), output_dim=64, input_resolution=32, width=rn_width)).eval()
        oclip.visual.load_state_dict({k[len("visual."):]: v for k, v in rn_sd.items()}, strict=False)
    else:
        pvit_cfg = pv.ViTConfig(32, 16, 128, 2, 64)
        oclip = seeded_init_(CLIPVisualOnly(OViTConfig(32, 16, 128, 2, 64)), seed=5).eval()
    for prm in list(ounet.parameters()) + list(oclip.parameters()):
        prm.requires_grad_(False)
    odiff = od.create_gaussian_diffusion(1000, "linear", respacing)
    pdiff = pgd.create_gaussian_diffusion(1000, "linear", respacing)
    g = th.Generator().manual_seed(99)
    targets = th.randn(P, 64, generator=g)
    weights = th.ones(P) / P
    kw = dict(clip_guidance_scale=1000.0, tv_scale=150.0, range_scale=50.0, sat_scale=sat_scale)
    kw.update(scales or {})
    lp_sd = init = olp = None
    if init_scale:
        from oracle import lpips as ol
        lp_sd = ol.seeded_state_dict()
        olp = ol.LPIPSVgg(lp_sd)
        init = th.rand(1, 3, image, image, generator=th.Generator().manual_seed(7)) * 2 - 1
    eng = pg.GuidedStepB200(prod_unet_cfg(ocfg), ounet.state_dict(), pvit_cfg, oclip.state_dict(), batch=B,
                            num_cutouts=cutn, max_prompts=P, use_magnitude=use_magnitude, device=device, conv_impl=conv_impl,
                            height=(hw or (image, image))[0], width=(hw or (image, image))[1],
                            rank=rank, world_size=world_size, use_graph=use_graph, vit_streams=vit_streams, cutn_variants=cutn_variants, lpips_sd=lp_sd, init_scale=init_scale, cutout_resize=cutout_resize, use_augs=use_augs, **kw)
    eng.set_targets(targets, weights)
    return dict(ounet=ounet, oclip=oclip, odiff=odiff, pdiff=pdiff, eng=eng, targets=targets, weights=weights, kw=kw, olp=olp, init=init,
                init_scale=init_scale, cutout_resize=cutout_resize,
                use_magnitude=use_magnitude, cutn=run_cutn or cutn, B=B, image=image, hw=hw or (image, image))


def oracle_step(ctx, mode, x, t_index, y, noise_seed, coords, fac_index, aug=None):
    """aug = (params [cutn, 20], noise [cutn, 4, B, 3, Smax, Smax]): the use_augs pipeline with the engine's draws (run the engine first)"""
    odiff = ctx["odiff"]
    cond = og.OracleCondFn(odiff, ctx["oclip"], ctx["targets"], ctx["weights"], cut_size=32, num_cutouts=ctx["cutn"],
                           use_magnitude=ctx["use_magnitude"], lpips_model=ctx.get("olp"), init_tensor=ctx.get("init"),
                           init_scale=ctx.get("init_scale", 0.0), cutout_resize=ctx.get("cutout_resize", "pool"), **ctx["kw"])
    cond.current_timestep = fac_index
    grabbed = {}

    def cond_fn(xx, tt, out, y=None):
        g = cond(xx, tt, out, y=y, coords=coords, **(dict(aug_params=aug[0], aug_noise=aug[1]) if aug is not None else {}))
        grabbed["g"] = g.detach().clone()
        return g

    t = th.full((x.shape[0],), t_index, dtype=th.long)
    th.manual_seed(noise_seed)
    fn = odiff.p_sample_with_grad if mode == "ancestral" else odiff.ddim_sample_with_grad
    out = fn(ctx["ounet"], x, t, clip_denoised=False, cond_fn=cond_fn, model_kwargs={"y": y})
    out["g"] = grabbed["g"]
    out["terms"] = cond.last_terms
    return out


def engine_step(ctx, mode, x, t_index, y, noise, coords, fac_index, runner=None, fused=False):
    eng, pdiff = ctx["eng"], ctx["pdiff"]
    sc = pdiff.scalar_table(t_index, fac_index, 0.0)
    eng.stage_step(sc, coords, pdiff.model_timestep(t_index), y)
    eng.img(eng.unet.x_in).copy_(x)
    eng.img(eng.noise).copy_(noise)
    cutn = len(coords) if eng.cutn else None
    if fused:
        eng.replay(mode, cutn)
    else:
        eng._run_all(mode, runner, cutn)
    if eng.device.type == "cuda":
        th.cuda.synchronize()
    return dict(sample=eng.img(eng.sample).float().cpu().clone(), pred_xstart=eng.img(eng.x0).float().cpu().clone(),
                g=eng.img(eng.g).float().cpu().clone(), losses={k: v.clone() for k, v in eng.losses().items()})


def make_inputs(ctx, seed=3):
    B, (H, W) = ctx["B"], ctx["hw"]
    g = th.Generator().manual_seed(seed)
    x = th.randn(B, 3, H, W, generator=g)
    y = th.randint(0, 10, (B,), generator=g)
    th.manual_seed(seed + 100)
    noise = th.randn_like(x)  # == what the oracle draws right after manual_seed(seed + 100)
    th.manual_seed(seed + 200)
    coords = og.MakeCutouts(32, ctx["cutn"])._generate_coords(H, W, ctx["cutn"])  # sic: (H, W) as (side_x, side_y), quirk B3
    return x, y, noise, seed + 100, coords


def compare(o, e):
    return dict(cos_g=cos(e["g"], o["g"]), rel_g=rel(e["g"], o["g"]), rel_x0=rel(e["pred_xstart"], o["pred_xstart"]),
                rel_sample=rel(e["sample"], o["sample"]))


def run_tiny_step_parity(device="cuda:0", mode="ancestral", t_index=14, runner_factory=None, **build_kw):
    ctx = build_tiny(device, **build_kw)
    if ctx["init"] is not None:
        ctx["eng"].set_init_image(ctx["init"], runner_factory(ctx["eng"]) if runner_factory else None)
    x, y, noise, nseed, coords = make_inputs(ctx)
    runner = runner_factory(ctx["eng"]) if runner_factory else None
    aug = None
    if build_kw.get("use_augs"):  # the engine draws this step's aug parameters / noise while staging; the oracle gets the same ones
        th.manual_seed(4242)
        e = engine_step(ctx, mode, x, t_index, y, noise, coords, fac_index=t_index, runner=runner)
        eng = ctx["eng"]
        aug = (eng.v(eng.aug_prm, (eng.cutn, 20)).float().cpu().clone(),
               eng.v(eng.aug_noise, (eng.cutn, 4, eng.B, 3, eng.aug_smax, eng.aug_smax)).float().cpu().clone())
        assert float(aug[0][:, 1:7].abs().sum()) > 0 and float(aug[1].abs().sum()) > 0
    o = oracle_step(ctx, mode, x, t_index, y, nseed, coords, fac_index=t_index, aug=aug)
    if aug is None:
        e = engine_step(ctx, mode, x, t_index, y, noise, coords, fac_index=t_index, runner=runner)
    res = compare(o, e)
    res["clip_loss_rel"] = abs(float(e["losses"]["clip"].sum()) - o["terms"]["clip"]) / (abs(o["terms"]["clip"]) + 1e-9)
    res["tv_loss_rel"] = abs(float(e["losses"]["tv"].sum()) - o["terms"]["tv"]) / (abs(o["terms"]["tv"]) + 1e-9)
    res["range_loss_rel"] = abs(float(e["losses"]["range"].sum()) - o["terms"]["range"]) / (abs(o["terms"]["range"]) + 1e-9)
    return res


def psnr(a, b, peak=None):
    """peak = the reference image's own range (a trained model's samples live in [-1, 1]; seeded random weights do not)"""
    peak = float(b.max() - b.min()) if peak is None else peak
    mse = float(((a.float() - b.float()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * float(np.log10(peak * peak / mse))


def run_tiny_chain(device="cuda:0", mode="ancestral", steps=None, runner_factory=None, fused=False, seed=11, **build_kw):
    """Free-running chain (SURVEY.md 8c protocol item 3): both sides start from the same x_T and receive the same per-step
    (class label, noise, cutout windows), but each side feeds ITS OWN sample back in -- nothing is teacher-forced after step 0.
    Returns the per-step relative drift and the PSNR of the final sample / pred_xstart against the oracle's."""
    ctx = build_tiny(device, **build_kw)
    T = ctx["pdiff"].num_timesteps
    steps = T if steps is None else steps
    runner = runner_factory(ctx["eng"]) if runner_factory else None
    g = th.Generator().manual_seed(seed)
    xo = th.randn(ctx["B"], 3, *ctx["hw"], generator=g)
    xe = xo.clone()
    drift = []
    o = e = None
    for k in range(steps):
        t_index = T - 1 - k
        y = th.randint(0, 10, (ctx["B"],), generator=g)
        th.manual_seed(seed + 1000 + k)
        noise = th.randn_like(xo)
        th.manual_seed(seed + 5000 + k)
        coords = og.MakeCutouts(32, ctx["cutn"])._generate_coords(*ctx["hw"], ctx["cutn"])
        o = oracle_step(ctx, mode, xo, t_index, y, seed + 1000 + k, coords, fac_index=t_index)
        e = engine_step(ctx, mode, xe, t_index, y, noise, coords, fac_index=t_index, runner=runner, fused=fused)
        xo, xe = o["sample"].detach(), e["sample"]
        drift.append(rel(xe, xo))
    return dict(drift=drift, psnr_sample=psnr(xe, xo), psnr_x0=psnr(e["pred_xstart"], o["pred_xstart"].detach()),
                finite=bool(th.isfinite(xe).all()), steps=steps)
