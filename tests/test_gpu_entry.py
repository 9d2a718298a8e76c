"""GPU: the reference's entry point surface: `clip_guided_diffusion(...)` is a generator of (batch_idx, png_path)
(cgd/cgd.py:19-55, 265-271) -- same keyword arguments, frames saved every `save_frequency` steps under
prefix_path/<prompt>/<batch idx>/<step>.png (cgd/script_util.py:93-101).  The numerics of the loop it drives are covered by
tests/test_gpu_chain.py and tests/test_loops_cpu.py; this checks the orchestration on the device."""
import os

import pytest
import torch as th

from clip_guided_diffusion_b200 import cgd
from clip_guided_diffusion_b200 import unet as pu
from clip_guided_diffusion_b200 import vit as pv
from clip_guided_diffusion_b200 import weights as pw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seeded_64():
    ucfg, vcfg = pu.config_for(64, True), pv.VIT_CONFIGS["ViT-B/32"]
    return pw.seeded_state_dict(pw.unet_param_shapes(ucfg), 1234), pw.seeded_state_dict(pw.vit_param_shapes(vcfg), 1235)


@pytest.mark.parametrize("respacing,batch,cutn,extra", [("25", 2, 4, {}), ("ddim25", 1, 16, dict(progressive_cutout=True, cached_cutouts=True))])
def test_clip_guided_diffusion_generator(tmp_path, monkeypatch, seeded_64, respacing, batch, cutn, extra):
    from PIL import Image
    monkeypatch.chdir(tmp_path)  # log_image also writes ./current.png like the reference
    usd, vsd = seeded_64
    tgt = th.randn(1, 512, generator=th.Generator().manual_seed(0))
    gen = cgd.clip_guided_diffusion(image_size=64, num_cutouts=cutn, prompts=["a test prompt"], batch_size=batch, timestep_respacing=respacing,
                                    skip_timesteps=22, save_frequency=2, prefix_path=tmp_path / "out", progress=False, seed=0,
                                    unet_state_dict=usd, clip_state_dict=vsd, target_embeds=tgt, weights=[1.0], **extra)
    got = list(gen)
    # 25 - 22 = 3 steps; frames at steps 0 and 2 (current_timestep never reaches -1 with skip_timesteps: quirk B2), every batch item
    assert [b for b, _ in got] == list(range(batch)) * 2
    for k, (b, path) in enumerate(got):
        step = 0 if k < batch else 2
        assert path.endswith(os.path.join("a_test_prompt", f"{b:02d}", f"{step:04d}.png")) and os.path.exists(path), path
        im = Image.open(path)
        assert im.size == (64, 64) and im.mode == "RGB"
    assert os.path.exists(tmp_path / "current.png")


def test_entry_rejects_what_the_reference_cannot_do(seeded_64):
    usd, vsd = seeded_64
    with pytest.raises(NotImplementedError):
        next(cgd.clip_guided_diffusion(image_size=64, image_prompts=["x.png"], unet_state_dict=usd, clip_state_dict=vsd,
                                       target_embeds=th.zeros(1, 512), weights=[1.0]))
    with pytest.raises(RuntimeError, match="sum to 0"):
        next(cgd.clip_guided_diffusion(image_size=64, unet_state_dict=usd, clip_state_dict=vsd, target_embeds=th.zeros(2, 512),
                                       weights=[1.0, -1.0]))
