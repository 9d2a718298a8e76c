"""GPU: every kernel of the guided step, op by op, against the PyTorch interpreter on identical inputs; then the
teacher-forced single-step parity against the oracle (eager and CUDA-graph replay)."""
import os

import pytest
import torch as th

from tests.gpu_harness import compare_ops
from tests.step_parity import build_tiny, compare, engine_step, make_inputs, oracle_step

pytestmark = pytest.mark.gpu
IMPL = int(os.environ.get("CGD_TEST_CONV_IMPL", "0"))

SEGS = [("unet_emb", "unet_bwd"), ("pmv", "cond"), ("cut_fwd", "sph"), ("vit_fwd", "vit_bwd"), ("sph", "cut_bwd"), ("vit_bwd", "vit_end"),
        ("cut_bwd", "guide"), ("guide", "final"), ("unet_bwd", "unet_end"), ("final", "upd_anc_g"), ("upd_anc_g", "upd_anc"),
        ("upd_ddim_g", "upd_ddim")]


@pytest.mark.parametrize("kw", [dict(image=32), dict(image=64, B=1, cutn=2, use_magnitude=True, sat_scale=20.0, new_order=True),
                                dict(image=64, B=2, cutn=3, cutout_resize="lanczos3"), dict(image=64, B=2, cutn=4, use_augs=True)],
                         ids=["b2_32px", "b1_64px_mag_sat_neworder", "b2_64px_resize_right", "b2_64px_use_augs"])
def test_every_op_matches_interpreter(kw):
    ctx = build_tiny("cuda", conv_impl=IMPL, **kw)
    eng = ctx["eng"]
    x, y, noise, nseed, coords = make_inputs(ctx)
    sc = ctx["pdiff"].scalar_table(14, 14, 0.0)
    th.manual_seed(77)  # use_augs: the aug parameters / noise fields are drawn while staging
    eng.stage_step(sc, coords, ctx["pdiff"].model_timestep(14), y)
    eng.img(eng.unet.x_in).copy_(x)
    eng.img(eng.noise).copy_(noise)
    th.cuda.synchronize()
    n, failures = compare_ops(eng.plan, SEGS)
    msg = "\n".join(str(f) for f in failures[:40])
    assert not failures, f"{len(failures)} of {n} ops differ from the interpreter:\n{msg}"


@pytest.mark.parametrize("streams", [1, 2], ids=["vit1stream", "vit2streams"])
@pytest.mark.parametrize("mode", ["ancestral", "ddim"])
@pytest.mark.parametrize("fused", [False, True], ids=["eager", "graph"])
def test_step_parity_vs_oracle(mode, fused, streams):
    ctx = build_tiny("cuda", conv_impl=IMPL, image=64, use_graph=True, cutn=4 if streams == 2 else 3, vit_streams=streams)
    x, y, noise, nseed, coords = make_inputs(ctx)
    o = oracle_step(ctx, mode, x, 14, y, nseed, coords, fac_index=14)
    e = engine_step(ctx, mode, x, 14, y, noise, coords, fac_index=14, fused=fused)
    if fused:  # replay a second time with fresh inputs staged: graph must pick up the new step data
        e = engine_step(ctx, mode, x, 14, y, noise, coords, fac_index=14, fused=True)
    res = compare(o, e)
    # tolerance: fp16 activations/weights vs the fp32 oracle, teacher-forced single step
    assert res["cos_g"] > 0.995 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res


@pytest.mark.parametrize("hw", [(64, 96), (96, 64)], ids=["64x96", "96x64"])
def test_non_square_step_vs_oracle(hw):
    """height_offset / width_offset (cgd/cgd.py:135): non-square x_t; cutout windows drawn with the reference's swapped sides are
    clipped at the border and still pooled to a square (quirk B3)"""
    ctx = build_tiny("cuda", conv_impl=IMPL, image=64, hw=hw, use_graph=True, B=2, cutn=5)
    x, y, noise, nseed, coords = make_inputs(ctx)
    o = oracle_step(ctx, "ddim", x, 14, y, nseed, coords, fac_index=14)
    e = engine_step(ctx, "ddim", x, 14, y, noise, coords, fac_index=14, fused=True)
    res = compare(o, e)
    assert res["cos_g"] > 0.995 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res
    with pytest.raises(RuntimeError, match="outside"):
        ctx["eng"].stage_step(ctx["pdiff"].scalar_table(14, 14, 0.0), [(0, hw[0], 32)] * 5, ctx["pdiff"].model_timestep(14), y)


@pytest.mark.parametrize("fused", [False, True], ids=["eager", "graph"])
def test_progressive_cutout_variant_vs_oracle(fused):
    """An engine built for cutout counts (2, 4, 8) -- progressive_cutout, cgd/cgd.py:167-175 -- runs a step with 4 cutouts
    (its own ViT activations / op ranges / CUDA graph, shared packed weights) and matches the oracle with the same 4 windows."""
    ctx = build_tiny("cuda", conv_impl=IMPL, image=64, use_graph=True, B=1, cutn=8, cutn_variants=(2, 4, 8), run_cutn=4)
    x, y, noise, nseed, coords = make_inputs(ctx)
    assert len(coords) == 4
    o = oracle_step(ctx, "ddim", x, 14, y, nseed, coords, fac_index=14)
    e = engine_step(ctx, "ddim", x, 14, y, noise, coords, fac_index=14, fused=fused)
    res = compare(o, e)
    assert res["cos_g"] > 0.995 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res


def test_step_with_lpips_init_loss_vs_oracle():
    """Whole guided step with an init image: ``+ lpips_vgg(x_in, init).sum() * init_scale`` inside the differentiated loss
    (cgd/cgd.py:220-224), CUDA-graph replay vs the oracle."""
    ctx = build_tiny("cuda", conv_impl=IMPL, image=64, use_graph=True, B=1, cutn=2, init_scale=1000.0)
    ctx["eng"].set_init_image(ctx["init"])
    x, y, noise, nseed, coords = make_inputs(ctx)
    o = oracle_step(ctx, "ddim", x, 14, y, nseed, coords, fac_index=14)
    e = engine_step(ctx, "ddim", x, 14, y, noise, coords, fac_index=14, fused=True)
    res = compare(o, e)
    assert res["cos_g"] > 0.995 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res
    assert abs(float(e["losses"]["init"].sum()) - o["terms"]["init"]) / abs(o["terms"]["init"]) < 3e-2


def test_step_with_resize_right_cutouts_vs_oracle():
    """Cutouts resampled by the ResizeRight lanczos3 tables (the mode north_star names) instead of pooling: graph replay vs the
    oracle whose MakeCutouts calls the pinned restatement of cgd/ResizeRight/resize_right.py."""
    ctx = build_tiny("cuda", conv_impl=IMPL, image=64, use_graph=True, B=2, cutn=3, cutout_resize="lanczos3")
    x, y, noise, nseed, coords = make_inputs(ctx)
    o = oracle_step(ctx, "ancestral", x, 14, y, nseed, coords, fac_index=14)
    e = engine_step(ctx, "ancestral", x, 14, y, noise, coords, fac_index=14, fused=True)
    res = compare(o, e)
    assert res["cos_g"] > 0.995 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res


@pytest.mark.parametrize("fused", [False, True], ids=["eager", "graph"])
def test_step_with_use_augs_vs_oracle(fused):
    """use_augs (cgd/modules.py:12-24, 62): random flip / affine / perspective / grayscale / noise on every cutout inside the cutout
    kernels.  The engine draws the parameters (CPU generator, torchvision's order) and noise fields (device generator) while staging
    the step; the oracle runs torchvision's own kernels with exactly those draws (oracle.guidance.apply_augs, pinned bit-exactly on
    the reference's MakeCutouts(use_augs=True) by tests/test_oracle.py)."""
    ctx = build_tiny("cuda", conv_impl=IMPL, image=64, use_graph=True, B=2, cutn=4, use_augs=True)
    eng = ctx["eng"]
    x, y, noise, nseed, coords = make_inputs(ctx)
    th.manual_seed(4242)
    e = engine_step(ctx, "ddim", x, 14, y, noise, coords, fac_index=14, fused=fused)
    aug = (eng.v(eng.aug_prm, (eng.cutn, 20)).float().cpu().clone(),
           eng.v(eng.aug_noise, (eng.cutn, 4, eng.B, 3, eng.aug_smax, eng.aug_smax)).float().cpu().clone())
    assert float(aug[1].abs().sum()) > 0
    o = oracle_step(ctx, "ddim", x, 14, y, nseed, coords, fac_index=14, aug=aug)
    res = compare(o, e)
    assert res["cos_g"] > 0.995 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res
