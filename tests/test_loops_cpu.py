"""CPU: the drop-in surface end to end -- `p_sample_loop_progressive` / `ddim_sample_loop_progressive` with a
`CondFnB200`, the call cgd/cgd.py:242-262 makes -- against the oracle's loops under the SAME torch seed: identical RNG draw order
(x_T, per-step class labels, ancestral noise before / DDIM noise after cond_fn, three CPU-generator draws per cutout), the
`current_timestep` bookkeeping (quirk B2), `skip_timesteps` + `init_image` (q_sample start), the fused-step path and the
segment path (p_mean_variance -> cond_fn -> update).  Kernels are interpreted (tests/plan_interp.py)."""
import pytest
import torch as th

from clip_guided_diffusion_b200 import guidance as pg
from oracle import guidance as og
from tests.plan_interp import Interp
from tests.step_parity import build_tiny, rel


def _interpreted(ctx):
    eng = ctx["eng"]
    it = Interp(eng.plan)
    eng.plan.run_range = lambda a, b, stream=None: it.run_range(a, b)
    eng.plan.run = lambda first=0, count=None, stream=None: it.run(first, len(eng.plan.ops) - first if count is None else count)
    return eng


@pytest.mark.parametrize("mode,fused,skip", [("ancestral", True, 0), ("ddim", True, 0), ("ancestral", False, 0), ("ddim", False, 21),
                                             ("ancestral", True, 20)])
def test_sampling_loops_follow_the_reference_draw_order(mode, fused, skip):
    B, cutn, image = 2, 2, 32
    ctx = build_tiny("cpu", B=B, cutn=cutn, image=image, use_magnitude=True)
    eng = _interpreted(ctx)
    if not fused:
        eng.can_fuse = lambda *a, **k: False
    pdiff, odiff = ctx["pdiff"], ctx["odiff"]
    T = pdiff.num_timesteps
    steps = 4 if skip == 0 else T - skip
    # the reference expands the init image to the batch before the loop (cgd/cgd.py:116-120)
    init = (th.rand(1, 3, image, image, generator=th.Generator().manual_seed(4)) * 2 - 1).expand(B, -1, -1, -1) if skip else None
    shape = (B, 3, image, image)
    y0 = th.zeros(B, dtype=th.long)

    cond = pg.CondFnB200(eng, pdiff, pg.MakeCutouts(32, cutn))
    ocond = og.OracleCondFn(odiff, ctx["oclip"], ctx["targets"], ctx["weights"], cut_size=32, num_cutouts=cutn, use_magnitude=True, **ctx["kw"])
    name = "p_sample_loop_progressive" if mode == "ancestral" else "ddim_sample_loop_progressive"
    kw = dict(clip_denoised=False, model_kwargs={"y": y0}, randomize_class=True, cond_fn_with_grad=True, skip_timesteps=skip, init_image=init,
              progress=False)

    th.manual_seed(0)
    outs_o = []
    for k, o in enumerate(getattr(odiff, name)(ctx["ounet"], shape, cond_fn=ocond, **kw)):
        outs_o.append((o["sample"].detach().clone(), o["pred_xstart"].detach().clone()))
        ocond.step_done()  # the reference's driver loop: cgd/cgd.py:265-267
        if k + 1 == steps:
            break
    rng_after_oracle = th.get_rng_state()

    th.manual_seed(0)
    outs_e = []
    for k, e in enumerate(getattr(pdiff, name)(eng.model, shape, cond_fn=cond, **kw)):
        outs_e.append((e["sample"].float().clone(), e["pred_xstart"].float().clone()))
        cond.step_done()
        if k + 1 == steps:
            break
    # same number of draws of the same kinds from the default generator on both sides
    assert th.equal(th.get_rng_state(), rng_after_oracle)
    assert cond.current_timestep == ocond.current_timestep == T - 1 - steps  # quirk B2: not advanced by skip_timesteps
    for k, ((so, xo), (se, xe)) in enumerate(zip(outs_o, outs_e)):
        assert rel(se, so) < 2e-2 and rel(xe, xo) < 2e-2, (k, rel(se, so), rel(xe, xo))


@pytest.mark.parametrize("fused", [True, False], ids=["fused_step", "segments"])
def test_reduce_clip_schedule_matches_the_reference_rule(fused):
    """reduce_clip (cgd/cgd.py:141-144, 157-164): the first 20 % of the steps are skipped through skip_timesteps, up to 70 % CLIP
    guidance runs on every 4th step and cond_fn returns zeros otherwise (the sampler then takes an unguided step).  Fused path: the
    skipped steps run the short op list (UNet forward -> p_mean_variance -> unguided update), on a GPU a second CUDA graph."""
    B, cutn, image = 1, 2, 32
    ctx = build_tiny("cpu", B=B, cutn=cutn, image=image, use_magnitude=True)
    eng = _interpreted(ctx)
    ran = []
    if fused:
        orig = eng.replay
        eng.replay = lambda mode, cutn=None, guided=True: (ran.append(guided), orig(mode, cutn, guided=guided))[1]
    else:
        eng.can_fuse = lambda *a, **k: False
    pdiff, odiff = ctx["pdiff"], ctx["odiff"]
    T = pdiff.num_timesteps
    skip = int(T * 0.2)
    cond = pg.CondFnB200(eng, pdiff, pg.MakeCutouts(32, cutn), reduce_clip=True)
    inner = og.OracleCondFn(odiff, ctx["oclip"], ctx["targets"], ctx["weights"], cut_size=32, num_cutouts=cutn, use_magnitude=True, **ctx["kw"])
    guided = []

    def ocond(x, t, out, y=None):  # the reference's rule, restated around the oracle's cond_fn
        pct = (T - inner.current_timestep) / T
        if pct < 0.7 and int((pct - 0.2) * T) % 4 != 0:
            guided.append(False)
            return th.zeros_like(x)
        guided.append(True)
        return inner(x, t, out, y=y)

    shape = (B, 3, image, image)
    kw = dict(clip_denoised=False, model_kwargs={"y": th.zeros(B, dtype=th.long)}, randomize_class=True, cond_fn_with_grad=True,
              skip_timesteps=skip, progress=False)
    th.manual_seed(1)
    outs_o = []
    for o in odiff.p_sample_loop_progressive(ctx["ounet"], shape, cond_fn=ocond, **kw):
        outs_o.append(o["sample"].detach().clone())
        inner.step_done()
    state = th.get_rng_state()
    th.manual_seed(1)
    outs_e = []
    for e in pdiff.p_sample_loop_progressive(eng.model, shape, cond_fn=cond, **kw):
        outs_e.append(e["sample"].float().clone())
        cond.step_done()
    assert th.equal(th.get_rng_state(), state)  # skipped steps draw no cutout windows on either side
    assert len(outs_e) == len(outs_o) == T - skip and 0 < sum(guided) < len(guided)
    if fused:
        assert ran == guided  # the engine replayed the unguided op list exactly on the steps the reference's rule skips
    worst = max(rel(a, b) for a, b in zip(outs_e, outs_o))
    assert worst < 3e-2, worst
