"""GPU, the larger BASELINE.json configurations (per-GPU shard): one guided step each, checked through size-independent
properties -- finite outputs, CUDA-graph replay == eager (bitwise) -- because the fp32 CPU oracle needs minutes per step here.
cfg4: image_size=512, ddim, 1 image per GPU, cutn=16, ViT-B/16.  cfg5: image_size=512, cutn=64, ViT-L/14, init image + LPIPS."""
import pytest
import torch as th

from clip_guided_diffusion_b200 import gaussian_diffusion as pgd
from clip_guided_diffusion_b200 import guidance as pg
from clip_guided_diffusion_b200 import unet as pu
from clip_guided_diffusion_b200 import vit as pv
from clip_guided_diffusion_b200 import weights as pw

pytestmark = pytest.mark.gpu


def _engine(clip_name, cutn, **kw):
    ucfg, vcfg = pu.config_for(512, True), pv.VIT_CONFIGS[clip_name]
    usd = pw.seeded_state_dict(pw.unet_param_shapes(ucfg), 1234)
    vsd = pw.seeded_state_dict(pw.vit_param_shapes(vcfg), 1235)
    eng = pg.GuidedStepB200(ucfg, usd, vcfg, vsd, batch=1, num_cutouts=cutn, device="cuda", **kw)
    th.manual_seed(0)
    eng.set_targets(th.randn(1, vcfg.output_dim), th.ones(1))
    diff = pgd.create_gaussian_diffusion(1000, "linear", "ddim250", rescale_timesteps=ucfg.rescale_timesteps)
    return eng, diff, vcfg


def _step(eng, diff, vcfg, mode, fused):
    th.manual_seed(3)
    x = th.randn(1, 3, 512, 512)
    y = th.tensor([5])
    noise = th.randn_like(x)
    th.manual_seed(9)
    coords = pg.MakeCutouts(vcfg.input_resolution, eng.cutn)._generate_coords(512, 512, eng.cutn)
    eng.stage_step(diff.scalar_table(200, 200, 0.0), coords, diff.model_timestep(200), y)
    eng.img(eng.unet.x_in).copy_(x)
    eng.img(eng.noise).copy_(noise)
    if fused:
        eng.replay(mode)
    else:
        eng._run_all(mode)
    th.cuda.synchronize()
    return {k: eng.img(b).clone() for k, b in (("sample", eng.sample), ("g", eng.g), ("x0", eng.x0))}


def test_cfg4_512_vit_b16_step():
    eng, diff, vcfg = _engine("ViT-B/16", 16)
    a = _step(eng, diff, vcfg, "ddim", fused=False)
    b = _step(eng, diff, vcfg, "ddim", fused=True)
    for k in a:
        assert th.isfinite(a[k]).all(), k
        assert th.equal(a[k], b[k]), f"graph replay differs from eager in {k}"
    assert float(a["g"].abs().max()) > 0


def test_cfg5_512_vit_l14_lpips_step():
    from oracle import lpips as ol  # seeded LPIPS weights in upstream key layout (test data only)
    eng, diff, vcfg = _engine("ViT-L/14", 64, lpips_sd=ol.seeded_state_dict(), init_scale=1000.0)
    th.manual_seed(4)
    eng.set_init_image((th.rand(1, 3, 512, 512) * 2 - 1).cuda())
    a = _step(eng, diff, vcfg, "ancestral", fused=False)
    b = _step(eng, diff, vcfg, "ancestral", fused=True)
    for k in a:
        assert th.isfinite(a[k]).all(), k
        assert th.equal(a[k], b[k]), f"graph replay differs from eager in {k}"
    losses = eng.losses()
    assert float(losses["init"].sum()) > 0 and th.isfinite(losses["clip"]).all()
