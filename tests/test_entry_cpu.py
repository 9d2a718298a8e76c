"""CPU: the orchestration of `clip_guided_diffusion(...)` (a generator of (batch_idx, png_path), cgd/cgd.py:19-55,
265-271) with the engine's kernels interpreted: the 64x64 checkpoint architecture, three steps.  The device check and the engine
class are patched (test seams only); the same call runs on the device in tests/test_gpu_entry.py."""
import os

import pytest
import torch as th

from clip_guided_diffusion_b200 import cgd
from clip_guided_diffusion_b200 import guidance as pg
from clip_guided_diffusion_b200 import unet as pu
from clip_guided_diffusion_b200 import vit as pv
from clip_guided_diffusion_b200 import weights as pw
from tests.plan_interp import Interp


class _InterpretedEngine(pg.GuidedStepB200):
    def __init__(self, *a, **k):
        k["device"] = "cpu"
        super().__init__(*a, **k)
        it = Interp(self.plan)
        self.plan.run_range = lambda a, b, stream=None: it.run_range(a, b)
        self.plan.run = lambda first=0, count=None, stream=None: it.run(first, len(self.plan.ops) - first if count is None else count)


def test_clip_guided_diffusion_generator_on_the_interpreter(tmp_path, monkeypatch):
    from PIL import Image
    with pytest.raises(RuntimeError, match="CUDA"):
        next(cgd.clip_guided_diffusion(image_size=64, device="cpu"))
    monkeypatch.setattr(cgd, "_require_cuda", lambda device: None)
    monkeypatch.setattr(cgd, "GuidedStepB200", _InterpretedEngine)
    monkeypatch.chdir(tmp_path)  # log_image also writes ./current.png like the reference
    ucfg, vcfg = pu.config_for(64, True), pv.ViTConfig(32, 16, 64, 1, 32)  # the real 64x64 UNet, a one-layer CLIP tower
    usd = pw.seeded_state_dict(pw.unet_param_shapes(ucfg), 1234)
    vsd = pw.seeded_state_dict(pw.vit_param_shapes(vcfg), 1235)
    tgt = th.randn(2, 32, generator=th.Generator().manual_seed(0))
    got = list(cgd.clip_guided_diffusion(image_size=64, num_cutouts=4, prompts=["a test prompt"], batch_size=1, timestep_respacing="25",
                                         skip_timesteps=22, save_frequency=2, prefix_path=tmp_path / "out", progress=False, seed=0,
                                         device="cpu", unet_state_dict=usd, clip_state_dict=vsd, target_embeds=tgt, weights=[2.0, 1.0],
                                         progressive_cutout=True))
    # 25 - 22 = 3 steps; frames at steps 0 and 2 (current_timestep never reaches -1 with skip_timesteps: quirk B2)
    assert [b for b, _ in got] == [0, 0]
    for (b, path), step in zip(got, (0, 2)):
        assert path.endswith(os.path.join("a_test_prompt", "00", f"{step:04d}.png")) and os.path.exists(path), path
        im = Image.open(path)
        assert im.size == (64, 64) and im.mode == "RGB"
    assert os.path.exists(tmp_path / "current.png")
    with pytest.raises(RuntimeError, match="sum to 0"):
        next(cgd.clip_guided_diffusion(image_size=64, device="cpu", unet_state_dict=usd, clip_state_dict=vsd, target_embeds=tgt, weights=[1.0, -1.0]))
    with pytest.raises(NotImplementedError):
        next(cgd.clip_guided_diffusion(image_size=64, device="cpu", image_prompts=["x.png"], unet_state_dict=usd, clip_state_dict=vsd,
                                       target_embeds=tgt, weights=[1.0, 1.0]))


def test_parse_prompt_and_log_image_known_answers(tmp_path, monkeypatch):
    """the reference's own known-answer tests for the entry's helpers (test.py:106-119, 178-194), plus the url branch and the
    file-name sanitising of cgd/script_util.py:60-67, 81-90"""
    monkeypatch.chdir(tmp_path)
    assert cgd.parse_prompt("Loose seal.:0.4") == ("Loose seal.", 0.4)
    assert cgd.parse_prompt("Loose seal.:-0.4") == ("Loose seal.", -0.4)
    assert cgd.parse_prompt("Loose seal.") == ("Loose seal.", 1.0)
    assert cgd.parse_prompt("https://a.b/c.png:2") == ("https://a.b/c.png", 2.0)
    assert cgd.parse_prompt("https://a.b/c.png") == ("https://a.b/c.png", 1.0)
    got = cgd.log_image(th.rand(3, 3, 3), str(tmp_path), ["a", "b", "c"], 1, 4)
    assert got == os.path.join(str(tmp_path), "a_b_c/04/0001.png") and os.path.exists(got)
    assert cgd.clean_and_combine_prompts("o", ["a cat: 0.5/x!", "b"], 0) == os.path.join("o", "a_cat_05x_b", "00")
    assert len(os.path.basename(os.path.dirname(cgd.clean_and_combine_prompts("o", ["x" * 400], 0)))) == 255
    assert cgd.DIFFUSION_FILENAMES[(False, 512)] == "512x512_diffusion_uncond_finetune_008100.pt"


def test_entry_helpers_match_the_reference_functions(tmp_path, monkeypatch):
    """parse_prompt / clean_and_combine_prompts / log_image against answers produced by executing the reference's own function bodies
    (tests/golden/make_golden_script_util.py): texts, weights, directory names, and the saved PNG pixel for pixel (torchvision's
    to_pil_image truncates `x * 255`, it does not round)"""
    import json
    import numpy as np
    from PIL import Image
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "script_util_golden.json")))
    for prompt, text, weight in gold["parse_prompt"]:
        assert cgd.parse_prompt(prompt) == (text, weight), prompt
    for texts, batch_idx, path in gold["clean_and_combine_prompts"]:
        assert cgd.clean_and_combine_prompts("o", texts, batch_idx) == path, texts
    monkeypatch.chdir(tmp_path)
    g = gold["log_image"]
    path = cgd.log_image(th.tensor(g["image"], dtype=th.float32), "out", ["a b", "c!"], 12, 3)
    assert path == g["path"] and os.path.exists("current.png") == g["current_png"]
    assert np.array_equal(np.asarray(Image.open(path)), np.asarray(g["pixels"], dtype=np.uint8))
    assert np.array_equal(np.asarray(Image.open("current.png")), np.asarray(g["pixels"], dtype=np.uint8))


def test_checkpoints_are_found_under_checkpoints_dir_by_the_reference_file_names(tmp_path):
    """`checkpoints_dir/<DIFFUSION_LOOKUP filename>` and `checkpoints_dir/clip/<CLIP_MODEL_URLS basename>` (cgd/script_util.py:18,
    cgd/clip_util.py:32-37): a plain state_dict file and a TorchScript archive (what OpenAI ships) both load; a missing file is a
    FileNotFoundError that names the path and the keyword to pass instead"""
    usd = {"a.weight": th.randn(3, 3)}
    th.save(usd, tmp_path / "256x256_diffusion_uncond.pt")
    got = cgd._load_unet_sd(256, False, str(tmp_path))
    assert th.equal(got["a.weight"], usd["a.weight"])
    with pytest.raises(FileNotFoundError, match="64x64_diffusion.pt"):
        cgd._load_unet_sd(64, True, str(tmp_path))
    with pytest.raises(ValueError, match="no published"):
        cgd._load_unet_sd(64, False, str(tmp_path))
    os.makedirs(tmp_path / "clip")
    th.save({"visual.proj": th.randn(4, 2)}, tmp_path / "clip" / "ViT-B-16.pt")  # plain dict: th.jit.load refuses it, th.load takes it
    assert tuple(cgd._load_clip_sd("ViT-B/16", str(tmp_path))["visual.proj"].shape) == (4, 2)

    class M(th.nn.Module):
        def __init__(self):
            super().__init__()
            self.visual = th.nn.Linear(2, 4)

        def forward(self, x):
            return self.visual(x)

    th.jit.script(M()).save(str(tmp_path / "clip" / "ViT-L-14-336px.pt"))  # a TorchScript archive
    assert set(cgd._load_clip_sd("ViT-L/14@336px", str(tmp_path))) == {"visual.weight", "visual.bias"}
    with pytest.raises(FileNotFoundError, match="RN50.pt"):
        cgd._load_clip_sd("RN50", str(tmp_path))
