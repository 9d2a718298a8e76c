"""CPU: the oracle's `cond_fn` restatement and the product's guidance schedules against the reference's OWN closure.

`tests/golden/cond_fn_golden.npz` holds what `cond_fn` of cgd/cgd.py:151-239 returned when its unmodified source was compiled and run
in the build container (tests/golden/make_golden_cond_fn.py: the nested function cut out with `ast`, its free variables supplied -- the
reference's real MakeCutouts / losses / CLIP_NORMALIZE, a deterministic stand-in for CLIP and for the UNet, both repeated below).
SURVEY.md 8c asks for exactly this kind of pin: "outputs of the reference itself run here"."""
import json
import os
import types

import numpy as np
import pytest
import torch as th
import torch.nn.functional as F

from oracle import diffusion as od
from oracle import guidance as og

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cond_fn_golden.npz")))
CASES = {  # the arguments of tests/golden/make_golden_cond_fn.py::CASES
    "plain": dict(B=1, P=1, hw=(48, 48), cutn=5, pow=1.0, T=25, t=10, sat=0.0, mag=False, seed=1),
    "late_step": dict(B=1, P=1, hw=(48, 48), cutn=3, pow=1.0, T=25, t=1, sat=0.0, mag=False, seed=2),
    "batch2_sat_magnitude": dict(B=2, P=1, hw=(48, 48), cutn=4, pow=1.0, T=25, t=20, sat=30.0, mag=True, seed=3),
    "three_prompts": dict(B=1, P=3, hw=(48, 48), cutn=4, pow=1.0, T=50, t=31, sat=0.0, mag=False, seed=4),
    "tall_image_cut_power": dict(B=1, P=1, hw=(48, 40), cutn=6, pow=0.5, T=25, t=15, sat=5.0, mag=True, seed=5),
}


class StubClip:  # same definition as in the golden script
    def __init__(self, cut_size, dim):
        k = 3 * (cut_size // 8) ** 2
        i, j = th.arange(k, dtype=th.float64).view(-1, 1), th.arange(dim, dtype=th.float64).view(1, -1)
        self.w = th.sin(0.37 * i + 0.11 * j * j + 0.5).float()

    def encode_image(self, img):
        return th.tanh(F.avg_pool2d(img, 8).flatten(1) @ self.w)


def stub_pred_xstart(x):  # same definition as in the golden script
    return th.tanh(1.5 * x.roll(1, -1)) * 1.2 + 0.05 * x ** 2


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_cond_fn_reproduces_the_reference_closure(name):
    c = CASES[name]
    x, want = th.from_numpy(GOLD[name + "_x"]), th.from_numpy(GOLD[name + "_grad"])
    diff = od.create_gaussian_diffusion(1000, "linear", str(c["T"]))
    # the oracle's own respaced table equals the one the golden run used (computed there from the DDPM formulas directly)
    assert abs(diff.sqrt_one_minus_alphas_cumprod[c["t"]] - float(GOLD[name + "_fac"][0])) < 1e-12
    cond = og.OracleCondFn(diff, StubClip(32, 16), th.from_numpy(GOLD[name + "_target"]), th.from_numpy(GOLD[name + "_weights"]), cut_size=32,
                           num_cutouts=c["cutn"], cutout_power=c["pow"], clip_guidance_scale=1000, tv_scale=150, range_scale=50, sat_scale=c["sat"],
                           use_magnitude=c["mag"])
    cond.current_timestep = c["t"]
    th.manual_seed(c["seed"])  # the windows come from the default CPU generator, three draws per cutout (cgd/modules.py:38-48)
    xr = x.clone().requires_grad_()
    got = cond(xr, th.full((c["B"],), c["t"]), {"pred_xstart": stub_pred_xstart(xr)})
    assert got.shape == want.shape
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5, name
    if c["mag"]:  # the clamp is over the WHOLE batch (cgd/cgd.py:229-232)
        assert float(got.square().mean().sqrt()) <= 0.05 * (1 + 1e-5)


def _stub_engine(counts):
    return types.SimpleNamespace(vits={c: None for c in counts}, cutn=max(counts))


def test_product_schedules_follow_the_reference_closure():
    """reduce_clip's every-4th-step rule and progressive_cutout's three counts (cgd/cgd.py:157-175), for every value of current_timestep
    of 25 / 50 / 250 / 1000-step chains and num_cutouts 6 / 8 / 16 / 32, as decided by the reference's closure itself"""
    from clip_guided_diffusion_b200 import guidance as pg
    sched = json.loads(bytes(GOLD["schedules_json"]).decode())
    assert len(sched) == 16
    for key, rows in sched.items():
        T, n = (int(v) for v in key.split("/"))
        diff = types.SimpleNamespace(num_timesteps=T)
        mk = types.SimpleNamespace(cutn=n)
        for (reduce_clip, progressive, cached), per_t in zip(((True, False, False), (False, True, True), (True, True, False)), rows):
            counts = pg.CondFnB200.progressive_counts(n) if progressive else (n,)
            cond = pg.CondFnB200(_stub_engine(counts), diff, mk, cached_cutouts=cached, reduce_clip=reduce_clip, progressive_cutout=progressive)
            assert cond.current_timestep == T - 1  # cgd/cgd.py:265
            for t, skipped, cutn, use_cache in per_t:
                cond.current_timestep = t
                assert cond.skips_guidance() == bool(skipped), (key, reduce_clip, progressive, t)
                if not skipped:
                    assert cond.current_cutn() == cutn and cond.cached_cutouts == bool(use_cache), (key, t)
