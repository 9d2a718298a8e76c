"""CPU: the whole guided step's op list (UNet fwd -> cutouts -> ViT fwd/bwd -> losses -> UNet dgrad -> update),
interpreted in PyTorch, against the oracle's p_sample_with_grad / ddim_sample_with_grad on identical inputs."""
import pytest

from tests.plan_interp import Interp
from tests.step_parity import run_tiny_step_parity


def _interp_runner(eng):
    it = Interp(eng.plan)
    return it.run_range


@pytest.mark.parametrize("mode,kw", [("ancestral", {}), ("ddim", {}), ("ancestral", dict(use_magnitude=True, sat_scale=30.0)),
                                     ("ddim", dict(B=1, P=2, cutn=2)), ("ddim", dict(B=2, cutn=4, vit_streams=2))])
def test_step_plan_matches_oracle(mode, kw):
    res = run_tiny_step_parity(device="cpu", mode=mode, runner_factory=_interp_runner, image=32, **kw)
    assert res["cos_g"] > 0.999, res
    assert res["rel_g"] < 5e-2 and res["rel_x0"] < 2e-2 and res["rel_sample"] < 2e-2, res
    assert res["clip_loss_rel"] < 2e-2 and res["tv_loss_rel"] < 1e-3 and res["range_loss_rel"] < 1e-2, res
