"""CPU: the whole guided step's op list (UNet fwd -> cutouts -> ViT fwd/bwd -> losses -> UNet dgrad -> update),
interpreted in PyTorch, against the oracle's p_sample_with_grad / ddim_sample_with_grad on identical inputs."""
import pytest

from tests.plan_interp import Interp
from tests.step_parity import run_tiny_step_parity


def _interp_runner(eng):
    it = Interp(eng.plan)
    return it.run_range


@pytest.mark.parametrize("mode,kw", [("ancestral", {}), ("ddim", {}), ("ancestral", dict(use_magnitude=True, sat_scale=30.0)),
                                     ("ddim", dict(B=1, P=2, cutn=2)), ("ddim", dict(B=2, cutn=4, vit_streams=2)),
                                     ("ancestral", dict(B=1, cutn=8, cutn_variants=(2, 4, 8), run_cutn=4)),
                                     ("ddim", dict(B=2, cutn=2, init_scale=1000.0)),
                                     ("ancestral", dict(B=2, cutn=3, image=64, cutout_resize="lanczos3")),
                                     # non-square (height_offset / width_offset, cgd/cgd.py:135): windows drawn with the reference's
                                     # swapped sides are clipped at the border (quirk B3) and still pooled to a square
                                     ("ddim", dict(B=1, cutn=6, image=32, hw=(32, 64))),
                                     ("ddim", dict(B=2, cutn=3, tower="rn")),  # CLIP ModifiedResNet tower instead of the ViT
                                     ("ddim", dict(B=1, cutn=2, tower="rn", rn_width=80)),  # RN50x4-style width: zero-padded 40 / 80 / 160-wide layers
                                     # use_augs (cgd/modules.py:12-24): flip / affine / perspective / grayscale / noise inside the cutout ops
                                     ("ddim", dict(B=2, cutn=4, image=64, use_augs=True)),
                                     ("ancestral", dict(B=1, cutn=6, image=32, hw=(32, 48), use_augs=True)),
                                     ("ancestral", dict(B=2, cutn=4, image=32, hw=(48, 32)))])
def test_step_plan_matches_oracle(mode, kw):
    res = run_tiny_step_parity(device="cpu", mode=mode, runner_factory=_interp_runner, **{"image": 32, **kw})
    assert res["cos_g"] > 0.999, res
    assert res["rel_g"] < 5e-2 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res
    assert res["clip_loss_rel"] < 2e-2 and res["tv_loss_rel"] < 1e-3 and res["range_loss_rel"] < 1e-2, res


def test_progressive_cutout_schedule():
    """CondFnB200 follows the reference's cutout-count schedule (cgd/cgd.py:167-175) and the CPU-generator draw order."""
    import torch as th
    from clip_guided_diffusion_b200 import guidance as pg

    class Eng:
        cutn = 16
        vits = {4: None, 8: None, 16: None}

    class Diff:
        num_timesteps = 100

    mk = pg.MakeCutouts(224, 16)
    cf = pg.CondFnB200(Eng(), Diff(), mk, progressive_cutout=True)
    seen = []
    for step in range(100):
        seen.append(cf.current_cutn())
        th.manual_seed(step)
        coords = cf.next_coords(256, 256)
        assert len(coords) == seen[-1]
        th.manual_seed(step)
        assert coords == mk._generate_coords(256, 256, seen[-1])  # same draws as the reference's MakeCutouts for that count
        cf.step_done()
    total = 100
    want = [4 if (total - t) / total < 0.3 else (8 if (total - t) / total < 0.7 else 16) for t in range(99, -1, -1)]
    assert seen == want
    with pytest.raises(ValueError):
        class Small(Eng):
            vits = {16: None}
        pg.CondFnB200(Small(), Diff(), mk, progressive_cutout=True)

    # num_cutouts < 16: the middle phase runs max(8, n // 2) = 8 cutouts, MORE than num_cutouts (the reference does the same);
    # the engine is then built for 8 with a 4-cutout variant (cgd.py)
    class Eng8:
        cutn = 8
        vits = {4: None, 8: None}
    mk4 = pg.MakeCutouts(224, 4)
    cf = pg.CondFnB200(Eng8(), Diff(), mk4, progressive_cutout=True)
    seen = []
    for step in range(100):
        seen.append(cf.current_cutn())
        assert len(cf.next_coords(256, 256)) == seen[-1]
        cf.step_done()
    assert seen == [4 if (100 - t) / 100 < 0.3 else (8 if (100 - t) / 100 < 0.7 else 4) for t in range(99, -1, -1)]
    # with cached_cutouts only num_cutouts windows exist: the reference crashes at .view([8, n, -1]); here a clear error
    mk4.cache_coordinates(256, 256)
    cf = pg.CondFnB200(Eng8(), Diff(), mk4, progressive_cutout=True, cached_cutouts=True)
    cf.current_timestep = 50
    with pytest.raises(RuntimeError, match="cached"):
        cf.next_coords(256, 256)


@pytest.mark.parametrize("mode", ["ancestral", "ddim"])
def test_short_chain_matches_oracle(mode):
    """free-running 6-step chain through the op-list interpreter: the host-side loop state (scalar tables, timestep map,
    fac index, buffer reuse between steps) stays consistent with the oracle's loop"""
    from tests.step_parity import run_tiny_chain
    res = run_tiny_chain(device="cpu", mode=mode, steps=6, runner_factory=_interp_runner, image=32, B=1, cutn=2, use_magnitude=True)
    assert res["finite"] and max(res["drift"]) < 2e-2 and res["psnr_sample"] > 40.0, res


def test_make_cutouts_surface():
    """cgd/modules.py:5-66: constructor, attributes, nn.Module-ness, coordinate cache, CPU-generator draw order; re-exported like the
    reference does (cgd/modules.py, cgd/clip_util.py:13)"""
    import torch as th
    from clip_guided_diffusion_b200 import clip_util, modules
    from clip_guided_diffusion_b200.guidance import MakeCutouts
    from oracle import guidance as og
    assert modules.MakeCutouts is MakeCutouts and clip_util.MakeCutouts is MakeCutouts
    mk = MakeCutouts(224, 16, cutout_size_power=0.5)
    assert isinstance(mk, th.nn.Module) and mk.to("cpu") is mk
    assert (mk.cut_size, mk.cutn, mk.cut_pow, mk.cached_coords) == (224, 16, 0.5, None) and hasattr(mk, "augs")
    th.manual_seed(0)
    mk.cache_coordinates(256, 320)
    th.manual_seed(0)
    ref = og.MakeCutouts(224, 16, 0.5)
    ref.cache_coordinates(256, 320)
    assert mk.cached_coords == ref.cached_coords and len(mk.cached_coords) == 16
    assert mk.coords_for(256, 320, use_cache=True, num_cutouts_override=4) == ref.cached_coords[:4]
    assert MakeCutouts(224, 16, use_augs=True).use_augs  # cgd/modules.py:12-24: the augmentations run inside the cutout kernels
    with pytest.raises(Exception):
        mk(th.zeros(1, 3, 256, 256))  # forward is the CUDA kernel: no CPU fallback
