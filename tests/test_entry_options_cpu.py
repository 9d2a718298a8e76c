"""CPU: every keyword of `clip_guided_diffusion(...)` that changes the per-step path (cgd/cgd.py:19-55), driven through the real entry
point with the engine's kernels interpreted and a two-level UNet of the 64 x 64 family (the real architecture runs in
tests/test_entry_cpu.py).  Checks the orchestration only: the run completes, yields (batch_idx, png_path) in the reference's order,
and the frames have the requested size; the arithmetic of each option is checked against the oracle in tests/test_step_cpu.py and on
the device."""
import os

import numpy as np
import pytest
import torch as th

from clip_guided_diffusion_b200 import cgd
from clip_guided_diffusion_b200 import unet as pu
from clip_guided_diffusion_b200 import vit as pv
from clip_guided_diffusion_b200 import weights as pw
from tests.test_entry_cpu import _InterpretedEngine

UCFG = pu.UNetConfig(image_size=64, model_channels=64, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(32,),
                     use_new_attention_order=True, noise_schedule="cosine")
VCFG = pv.ViTConfig(32, 16, 64, 1, 32)

# keyword overrides -> (frames yielded, PIL size); base call: "25" steps, skip 22 (3 steps), save_frequency 1, batch 1, 2 cutouts
CASES = {
    "ancestral": (dict(), 3, (64, 64)),
    "ddim": (dict(timestep_respacing="ddim25"), 3, (64, 64)),
    "reduce_clip": (dict(reduce_clip=True), 3, (64, 64)),
    # reduce_clip with skip_timesteps == 0 skips the first 20 % of the chain (cgd/cgd.py:141-144): 25 - 5 = 20 steps
    "reduce_clip_skips_a_fifth": (dict(reduce_clip=True, timestep_respacing="ddim25", skip_timesteps=0), 20, (64, 64)),
    "cached_cutouts": (dict(cached_cutouts=True), 3, (64, 64)),
    "progressive_cutout": (dict(progressive_cutout=True, num_cutouts=8), 3, (64, 64)),
    "use_augs": (dict(use_augs=True), 3, (64, 64)),
    "taller": (dict(height_offset=32), 3, (64, 96)),  # PIL size is (width, height)
    "init_image": (dict(init_image="init.png", skip_timesteps=20), 5, (64, 64)),
    "lpips_init_loss": (dict(init_image="init.png", skip_timesteps=22, init_scale=1000, lpips_state_dict="seeded"), 3, (64, 64)),
    "sat_and_magnitude": (dict(sat_scale=100.0, use_magnitude=True), 3, (64, 64)),
    "batch2_fixed_class": (dict(batch_size=2, randomize_class=False), 6, (64, 64)),
    "unconditional": (dict(class_cond=False), 3, (64, 64)),
    "cosine_schedule_cutout_power": (dict(noise_schedule="cosine", cutout_power=0.5, seed=5), 3, (64, 64)),
}


def _call(tmp_path, monkeypatch, kw):
    from PIL import Image
    ucfg = pu.UNetConfig(**{**UCFG.__dict__, "class_cond": kw.get("class_cond", True)})
    monkeypatch.setattr(cgd, "_require_cuda", lambda device: None)
    monkeypatch.setattr(cgd, "GuidedStepB200", _InterpretedEngine)
    monkeypatch.setattr(cgd, "config_for", lambda image_size, class_cond=True: ucfg)
    monkeypatch.chdir(tmp_path)
    Image.fromarray((np.random.default_rng(0).random((70, 90, 3)) * 255).astype("uint8")).save("init.png")
    base = dict(image_size=64, num_cutouts=2, prompts=["m"], batch_size=1, timestep_respacing="25", skip_timesteps=22, save_frequency=1,
                prefix_path=tmp_path / "out", progress=False, seed=0, device="cpu", unet_state_dict=pw.seeded_state_dict(pw.unet_param_shapes(ucfg), 1234),
                clip_state_dict=pw.seeded_state_dict(pw.vit_param_shapes(VCFG), 1235), target_embeds=th.randn(1, 32, generator=th.Generator().manual_seed(0)))
    base.update(kw)
    if base.get("lpips_state_dict") == "seeded":
        base["lpips_state_dict"] = pw.seeded_lpips_state_dict()
    return list(cgd.clip_guided_diffusion(**base))


@pytest.mark.parametrize("name", list(CASES))
def test_entry_option(name, tmp_path, monkeypatch):
    from PIL import Image
    kw, n_frames, size = CASES[name]
    got = _call(tmp_path, monkeypatch, kw)
    B = kw.get("batch_size", 1)
    assert len(got) == n_frames and [b for b, _ in got] == list(range(B)) * (n_frames // B)
    for _, path in got:
        assert os.path.exists(path)
    im = Image.open(got[-1][1])
    assert im.size == size and np.isfinite(np.asarray(im, dtype=np.float32)).all()


def test_wide_image_fails_like_the_reference(tmp_path, monkeypatch):
    """MakeCutouts reads (side_x, side_y) = input.shape[2:4], i.e. swapped (cgd/modules.py:52): on a wide image a window can start
    below the last row, the crop is empty and adaptive_avg_pool2d raises -- the replacement raises too instead of reading out of bounds"""
    with pytest.raises(RuntimeError, match="outside"):
        _call(tmp_path, monkeypatch, dict(width_offset=64, timestep_respacing="ddim25"))


def test_init_scale_without_lpips_weights_is_a_clear_error(tmp_path, monkeypatch):
    with pytest.raises(RuntimeError, match="LPIPS"):
        _call(tmp_path, monkeypatch, dict(init_image="init.png", skip_timesteps=20, init_scale=1000))


def test_offsets_that_break_the_skip_connections_are_a_clear_error(tmp_path, monkeypatch):
    """image_size + offset must stay a multiple of 2^(levels - 1) (the reference dies in th.cat of the first skip connection)"""
    with pytest.raises(ValueError, match="multiples of 2"):
        _call(tmp_path, monkeypatch, dict(height_offset=1))
    from clip_guided_diffusion_b200 import guidance as pg
    with pytest.raises(ValueError, match="multiples of 32 for this 6-level UNet"):
        pg.GuidedStepB200(pu.config_for(256, True), {}, None, None, batch=1, height=264, device="cpu")


def _driver_golden():
    import json
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "driver_loop_golden.json")))


@pytest.mark.parametrize("k", [i for i, g in enumerate(_driver_golden()) if g["respacing"] == "25" or g["skip"] >= 20])
def test_saved_frames_follow_the_reference_driver_loop(k, tmp_path, monkeypatch):
    """which (batch_idx, step) frames are yielded, against the reference's own loop statements executed on stand-ins
    (tests/golden/make_golden_driver_loop.py: cgd/cgd.py:241-271 cut out with `ast`) -- including quirk B2: `current_timestep` starts at
    num_timesteps - 1 whatever `skip_timesteps` is, so with skipped steps it never reaches -1 and the final frame is only saved when
    its step index happens to be a multiple of `save_frequency`"""
    g = _driver_golden()[k]
    got = _call(tmp_path, monkeypatch, dict(timestep_respacing=g["respacing"], skip_timesteps=g["skip"], save_frequency=g["save_frequency"],
                                            batch_size=g["batch"]))
    steps = [[b, int(os.path.basename(p).split(".")[0])] for b, p in got]
    assert steps == g["yields"], (g, steps)


def test_noise_schedule_is_validated_like_load_guided_diffusion(tmp_path, monkeypatch):
    with pytest.raises(ValueError, match="linear_or_cosine"):  # cgd/script_util.py:302-303
        _call(tmp_path, monkeypatch, dict(noise_schedule="sqrt"))
