"""GPU: free-running chains against the fp32 oracle (SURVEY.md 8c protocol item 3).  Both sides get the same x_T and the same
per-step class label / noise / cutout windows, but each feeds its OWN sample back in: fp16 error may accumulate over the 25
steps and is bounded here.  `use_magnitude` is on as the reference sets it for 64x64 models (cgd/cgd.py:72-74); with seeded random
weights it is also what keeps the chain itself bounded -- without the RMS clamp the ORACLE's own samples grow to 1e8 within two
steps (pred_xstart ~ 600 at t = T-1 feeds tv / range gradients of 5e4), which is a property of random weights, not of either
implementation."""
import pytest
import torch as th

from tests.step_parity import psnr, rel, run_tiny_chain

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["ancestral", "ddim"])
def test_tiny_chain_25_steps_vs_oracle(mode):
    res = run_tiny_chain(device="cuda:0", mode=mode, fused=True, use_graph=True, B=2, cutn=3, image=64, use_magnitude=True)
    assert res["finite"] and res["steps"] == 25, res
    assert max(res["drift"]) < 3e-2 and res["psnr_sample"] > 40.0 and res["psnr_x0"] > 40.0, res


def test_cfg1_chain_vs_oracle():
    """BASELINE configs[0] end to end: the 64x64 checkpoint architecture (296 M parameters), respace 25, batch 1, 4 cutouts,
    ViT-B/32, all 25 ancestral steps, oracle on the host CPU."""
    from clip_guided_diffusion_b200 import gaussian_diffusion as pgd
    from clip_guided_diffusion_b200 import guidance as pg
    from clip_guided_diffusion_b200 import unet as pu
    from clip_guided_diffusion_b200 import vit as pv
    from clip_guided_diffusion_b200 import weights as pw
    from oracle import diffusion as od
    from oracle import guidance as og
    from oracle.clip_vit import VIT_CONFIGS, CLIPVisualOnly
    from oracle.unet import UNetModel, config_for
    ucfg, vcfg = pu.config_for(64, True), pv.VIT_CONFIGS["ViT-B/32"]
    usd = pw.seeded_state_dict(pw.unet_param_shapes(ucfg), 1234)
    vsd = pw.seeded_state_dict(pw.vit_param_shapes(vcfg), 1235)
    ounet = UNetModel(config_for(64, True)).eval()
    ounet.load_state_dict(usd)
    oclip = CLIPVisualOnly(VIT_CONFIGS["ViT-B/32"]).eval()
    oclip.load_state_dict(vsd)
    for p in list(ounet.parameters()) + list(oclip.parameters()):
        p.requires_grad_(False)
    odiff = od.create_gaussian_diffusion(1000, "linear", "25")
    pdiff = pgd.create_gaussian_diffusion(1000, "linear", "25")
    th.manual_seed(0)
    tgt = th.randn(1, 512)
    eng = pg.GuidedStepB200(ucfg, usd, vcfg, vsd, batch=1, num_cutouts=4, use_magnitude=True, device="cuda")
    eng.set_targets(tgt, th.ones(1))
    coords = [(0, 0, 64)] * 4  # 64^2: min == max == 64 -> whole-image windows
    cond = og.OracleCondFn(odiff, oclip, tgt, th.ones(1), cut_size=224, num_cutouts=4, use_magnitude=True)
    g = th.Generator().manual_seed(3)
    xo = th.randn(1, 3, 64, 64, generator=g)
    xe = xo.clone()
    T = pdiff.num_timesteps
    drift = []
    for k in range(T):
        t_index = T - 1 - k
        y = th.randint(0, 1000, (1,), generator=g)
        cond.current_timestep = t_index
        th.manual_seed(100 + k)
        o = odiff.p_sample_with_grad(ounet, xo, th.tensor([t_index]), clip_denoised=False,
                                     cond_fn=lambda xx, tt, out, y=None: cond(xx, tt, out, y=y, coords=coords), model_kwargs={"y": y})
        th.manual_seed(100 + k)
        noise = th.randn_like(xo)
        eng.stage_step(pdiff.scalar_table(t_index, t_index, 0.0), coords, pdiff.model_timestep(t_index), y)
        eng.img(eng.unet.x_in).copy_(xe)
        eng.img(eng.noise).copy_(noise)
        eng.replay("ancestral")
        th.cuda.synchronize()
        xo, xe = o["sample"].detach(), eng.img(eng.sample).float().cpu().clone()
        drift.append(rel(xe, xo))
    res = dict(drift_max=max(drift), drift_last=drift[-1], psnr=psnr(xe, xo), finite=bool(th.isfinite(xe).all()))
    assert res["finite"] and res["drift_max"] < 5e-2 and res["psnr"] > 35.0, res


@pytest.mark.parametrize("resize", ["pool", "lanczos3"])
def test_loop_without_per_step_sync_equals_synced_run(resize):
    """The host runs ahead of the device between saved frames (default save_frequency 25, cgd/cgd.py:41): the per-step pinned
    staging (scalars, t, classes, cutout windows, ResizeRight tables) must not be rewritten before its copy executed.  The device
    is held busy while the host enqueues the whole chain, so an unguarded staging buffer would feed the last step's scalars to
    every step; the result must equal the run that synchronises after every step."""
    from clip_guided_diffusion_b200 import guidance as pg
    from tests.step_parity import build_tiny
    ctx = build_tiny("cuda:0", B=2, cutn=3, image=64, use_magnitude=True, use_graph=True, cutout_resize=resize)
    eng, pdiff = ctx["eng"], ctx["pdiff"]
    mc = pg.MakeCutouts(32, 3)
    finals = []
    for sync in (True, False):
        th.manual_seed(5)
        cond = pg.CondFnB200(eng, pdiff, mc)
        x_T = eng.draw_initial_noise()
        if not sync:
            th.cuda.synchronize()
            th.cuda._sleep(int(1.5e9))  # ~1 s at 1.5 GHz: every step of the chain is enqueued behind it
        loop = pdiff.ddim_sample_loop_progressive(eng.model, eng.shape, noise=x_T, clip_denoised=False, cond_fn=cond,
                                                  model_kwargs={"y": th.zeros(2, dtype=th.long, device="cuda")}, randomize_class=True,
                                                  cond_fn_with_grad=True)
        n, out = 0, None
        for out in loop:
            cond.step_done()
            n += 1
            if sync:
                th.cuda.synchronize()
        th.cuda.synchronize()
        assert n == 25
        finals.append((out["sample"].float().cpu(), out["pred_xstart"].float().cpu()))
    (s0, p0), (s1, p1) = finals
    assert th.isfinite(s0).all() and rel(s1, s0) < 1e-5 and rel(p1, p0) < 1e-5, (rel(s1, s0), rel(p1, p0))


@pytest.mark.parametrize("mode", ["ancestral", "ddim"])
def test_reduce_clip_fused_graphs_equal_segment_path(mode):
    """reduce_clip on the device (cgd/cgd.py:141-144, 157-164): the fused path replays two CUDA graphs -- the full guided step and the
    short unguided one (UNet forward -> p_mean_variance -> update) on the steps the rule skips -- and must reproduce the segment path
    (p_mean_variance -> cond_fn returning zeros / the gradient -> update, eager launches) over the whole chain."""
    from clip_guided_diffusion_b200 import guidance as pg
    from tests.step_parity import build_tiny
    ctx = build_tiny("cuda:0", B=2, cutn=3, image=64, use_magnitude=True, use_graph=True)
    eng, pdiff = ctx["eng"], ctx["pdiff"]
    T = pdiff.num_timesteps
    skip = int(T * 0.2)
    loop = pdiff.p_sample_loop_progressive if mode == "ancestral" else pdiff.ddim_sample_loop_progressive
    finals, kinds = [], []
    for fused in (True, False):
        th.manual_seed(7)
        cond = pg.CondFnB200(eng, pdiff, pg.MakeCutouts(32, 3), reduce_clip=True)
        if not fused:
            eng.can_fuse = lambda *a, **k: False
        else:
            orig = eng.replay
            eng.replay = lambda m, cutn=None, guided=True: (kinds.append(guided), orig(m, cutn, guided=guided))[1]
        out, n = None, 0
        for out in loop(eng.model, eng.shape, clip_denoised=False, cond_fn=cond, model_kwargs={"y": th.zeros(2, dtype=th.long, device="cuda")},
                        randomize_class=True, cond_fn_with_grad=True, skip_timesteps=skip):
            cond.step_done()
            n += 1
        th.cuda.synchronize()
        assert n == T - skip
        finals.append((out["sample"].float().cpu(), out["pred_xstart"].float().cpu()))
    assert 0 < sum(kinds) < len(kinds) == T - skip and (("ancestral", "unguided") in eng._graphs or ("ddim", "unguided") in eng._graphs)
    (s0, p0), (s1, p1) = finals
    assert th.isfinite(s0).all() and rel(s0, s1) < 1e-3 and rel(p0, p1) < 1e-3, (rel(s0, s1), rel(p0, p1))
