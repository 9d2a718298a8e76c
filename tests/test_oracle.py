"""CPU tests: the oracle against the reference-generated golden vectors (in-tree pieces) and the
published structural constants (third-party pieces)."""
import numpy as np
import pytest
import torch as th

from oracle import guidance as og
from oracle import diffusion as od
from oracle.unet import UNetModel, config_for, tiny_config, seeded_init_
from oracle.clip_vit import VIT_CONFIGS, CLIPVisualOnly, ViTConfig


def T(a):
    return th.from_numpy(np.asarray(a))


def test_losses_match_reference_goldens(golden):
    g = golden
    assert np.allclose(og.spherical_dist_loss(T(g["sph_small_x"]), T(g["sph_small_y"])).numpy(), g["sph_small"], atol=1e-7)
    assert abs(float(g["sph_small"][0]) - 0.49663094) < 1e-6
    img = T(g["img_small"])
    assert np.allclose(og.range_loss(img).numpy(), g["range_small"], atol=1e-7)
    assert np.allclose(og.tv_loss(img).numpy(), g["tv_small"], atol=1e-6)
    assert np.allclose(g["range_small"], [0.64526457, 0.58686501], atol=1e-6)
    assert np.allclose(g["tv_small"], [7.11696005, 7.48520803], atol=1e-5)


def test_loss_gradients_match_reference_goldens(golden):
    g = golden
    img = T(g["img_med"]).requires_grad_()
    r, t = og.range_loss(img), og.tv_loss(img)
    assert np.allclose(r.detach().numpy(), g["range_med"], atol=1e-7)
    assert np.allclose(t.detach().numpy(), g["tv_med"], atol=1e-6)
    assert np.allclose(th.autograd.grad(r.sum(), img, retain_graph=True)[0].numpy(), g["range_med_grad"], atol=1e-8)
    assert np.allclose(th.autograd.grad(t.sum(), img)[0].numpy(), g["tv_med_grad"], atol=1e-7)
    emb = T(g["sph_emb"]).requires_grad_()
    d = og.spherical_dist_loss(emb.unsqueeze(0), T(g["sph_tgt"]).unsqueeze(0))
    assert np.allclose(d.detach().numpy(), g["sph_med"], atol=1e-6)
    assert np.allclose(th.autograd.grad(d.sum(), emb)[0].numpy(), g["sph_med_grad"], atol=1e-6)


def test_cutout_coordinate_law(golden):
    th.manual_seed(0)
    c = og.MakeCutouts(224, 4, 1.0)._generate_coords(256, 256, 4)
    assert np.array_equal(np.array(c), golden["coords_256_seed0"])
    assert c == [(9, 11, 239), (26, 6, 228), (1, 1, 239), (12, 11, 244)]
    th.manual_seed(0)
    c = og.MakeCutouts(224, 6, 0.5)._generate_coords(512, 512, 6)
    assert np.array_equal(np.array(c), golden["coords_512_pow05_seed0"])
    # 64^2 checkpoints: min == max == 64 -> whole-image window
    assert og.MakeCutouts(224, 3)._generate_coords(64, 64, 3) == [(0, 0, 64)] * 3


def test_cutouts_values_and_grad(golden):
    g = golden
    src = T(g["cut_src"]).requires_grad_()
    coords = [tuple(int(v) for v in r) for r in g["cut_coords"]]
    cut = og.MakeCutouts(32, 3)(src, coords=coords)
    assert cut.shape == (6, 3, 32, 32)
    assert np.allclose(cut.detach().numpy(), g["cut_out"], atol=1e-6)
    grad = th.autograd.grad((cut * T(g["cut_wgt"])).sum(), src)[0]
    assert np.allclose(grad.numpy(), g["cut_grad"], atol=1e-5)
    th.manual_seed(3)  # same CPU-RNG draws as the reference run
    cut2 = og.MakeCutouts(32, 3)(src.detach())
    assert np.allclose(cut2.numpy(), g["cut_out"], atol=1e-6)
    up = og.MakeCutouts(56, 2)(T(g["cut_up_src"]), coords=[tuple(int(v) for v in r) for r in g["cut_up_coords"]])
    assert np.allclose(up.numpy(), g["cut_up_out"], atol=1e-6)


def test_reference_test_make_cutouts_shape():
    # /root/reference/test.py:233-249: [1,3,512,512], cut 224, cutn 8, pow 0.5 -> [8,3,224,224]
    out = og.MakeCutouts(224, 8, 0.5)(th.rand(1, 3, 512, 512))
    assert out.shape == (8, 3, 224, 224)


def test_schedule_constants():
    # SURVEY Appendix E (derived from the published formulas)
    d = od.create_gaussian_diffusion(1000, "linear", "25")
    assert d.timestep_map[:4] == [0, 42, 83, 125] and d.timestep_map[-3:] == [916, 957, 999]
    assert abs(d.betas[0] - 1e-4) < 1e-12 and abs(d.betas[24] - 0.5643942) < 1e-6
    s = d.sqrt_one_minus_alphas_cumprod
    assert abs(s[0] - 0.010000) < 1e-6 and abs(s[12] - 0.960314) < 1e-5 and abs(s[24] - 0.99997982) < 1e-7
    assert abs(d.posterior_variance[1] - 9.95564e-5) < 1e-9
    d = od.create_gaussian_diffusion(1000, "linear", "ddim250")
    assert d.timestep_map[:3] == [0, 4, 8] and d.timestep_map[-1] == 996 and d.num_timesteps == 250
    assert abs(d.betas[249] - 7.729432e-2) < 1e-7 and abs(d.sqrt_one_minus_alphas_cumprod[249] - 0.99997856) < 1e-7
    d = od.create_gaussian_diffusion(1000, "linear", "1000")
    assert abs(d.betas[999] - 0.02) < 1e-12 and abs(d.posterior_variance[1] - 5.453188e-5) < 1e-10
    d = od.create_gaussian_diffusion(1000, "cosine", "1000")
    assert abs(d.betas[0] - 4.128422e-5) < 1e-9 and d.betas[999] == 0.999
    assert abs(d.sqrt_one_minus_alphas_cumprod[500] - 0.712541) < 1e-5


def _nparams(m):
    return sum(p.numel() for p in m.parameters())


@pytest.mark.parametrize("size,cond,expect", [(64, True, 295.9e6), (256, True, 553.8e6), (256, False, 552.8e6),
                                              (512, True, 559.0e6), (128, True, 421.5e6)])
def test_unet_param_counts(size, cond, expect):
    with th.device("meta"):
        m = UNetModel(config_for(size, cond))
    assert abs(_nparams(m) - expect) < 0.06e6, _nparams(m)


@pytest.mark.parametrize("name,expect", [("ViT-B/32", 87.8e6), ("ViT-B/16", 86.2e6), ("ViT-L/14", 304.0e6)])
def test_vit_param_counts(name, expect):
    with th.device("meta"):
        m = CLIPVisualOnly(VIT_CONFIGS[name])
    assert abs(_nparams(m) - expect) < 0.06e6, _nparams(m)


def test_unet_state_dict_keys():
    with th.device("meta"):
        m = UNetModel(config_for(256, True))
    keys = set(m.state_dict().keys())
    for k in ["time_embed.0.weight", "time_embed.2.bias", "label_emb.weight", "input_blocks.0.0.weight",
              "input_blocks.1.0.in_layers.0.weight", "input_blocks.1.0.in_layers.2.bias",
              "input_blocks.1.0.emb_layers.1.weight", "input_blocks.1.0.out_layers.3.weight",
              "input_blocks.3.0.in_layers.2.weight", "middle_block.1.qkv.weight", "middle_block.1.proj_out.bias",
              "middle_block.1.norm.weight", "output_blocks.0.0.skip_connection.weight", "out.0.weight", "out.2.bias"]:
        assert k in keys, k
    assert m.state_dict()["middle_block.1.qkv.weight"].shape == (3072, 1024, 1)


def test_tiny_unet_and_step_run():
    cfg = tiny_config()
    m = seeded_init_(UNetModel(cfg)).eval()
    x = th.randn(2, 3, 32, 32)
    out = m(x, th.tensor([3.0, 500.0]), th.tensor([1, 2]))
    assert out.shape == (2, 6, 32, 32) and th.isfinite(out).all() and out.std() > 1e-3
    vit = seeded_init_(CLIPVisualOnly(ViTConfig(64, 32, 128, 2, 64)), seed=5).eval()
    diff = od.create_gaussian_diffusion(1000, "linear", "25")
    tgt = th.nn.functional.normalize(th.randn(1, 64), dim=-1)
    cond = og.OracleCondFn(diff, vit, tgt, th.tensor([1.0]), cut_size=64, num_cutouts=2)
    th.manual_seed(0)
    gen = diff.p_sample_loop_progressive(m, (2, 3, 32, 32), clip_denoised=False, model_kwargs={"y": th.zeros(2, dtype=th.long)},
                                         cond_fn=cond, randomize_class=True, cond_fn_with_grad=True)
    o = next(gen)
    assert o["sample"].shape == (2, 3, 32, 32) and th.isfinite(o["sample"]).all()
    gen2 = diff.ddim_sample_loop_progressive(m, (1, 3, 32, 32), clip_denoised=False, model_kwargs={"y": th.zeros(1, dtype=th.long)},
                                             cond_fn=cond, randomize_class=True, cond_fn_with_grad=True)
    o = next(gen2)
    assert th.isfinite(o["sample"]).all()


# ------------------------------------------------------------------ ResizeRight (cutout resize mode named by north_star)
RR_CASES = ["down_237_224", "down_64_56", "down_60_32", "down_73_32", "same_32", "up_16_56", "up_20_32"]


@pytest.mark.parametrize("name", RR_CASES)
def test_resize_right_restatement_pinned_to_reference(name):
    """oracle/resize_right.py against vectors produced by the reference's own cgd/ResizeRight (tests/golden/make_golden_resize.py):
    forward bit-exact, autograd gradient to float rounding."""
    import os
    import numpy as np
    from oracle.resize_right import resize_lanczos3
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize_right_golden.npz"))
    S, O, C = (int(v) for v in d[name + "_shape"])
    x = th.from_numpy(d[name + "_x"]).requires_grad_()
    y = resize_lanczos3(x, (O, O))
    assert th.equal(y.detach(), th.from_numpy(d[name + "_y"]))
    (g,) = th.autograd.grad((y * th.from_numpy(d[name + "_w"])).sum(), x)
    assert th.allclose(g, th.from_numpy(d[name + "_g"]), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", RR_CASES)
def test_resize_tables_reproduce_reference(name):
    """the product's host tables (clip_guided_diffusion_b200/resize_right.py) applied separably == the reference's output"""
    import os
    import numpy as np
    from clip_guided_diffusion_b200 import resize_right as rr
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize_right_golden.npz"))
    S, O, C = (int(v) for v in d[name + "_shape"])
    left, w, T = rr.tables(S, O)
    x = th.from_numpy(d[name + "_x"])[0]
    idx = left[:, None].long() + th.arange(rr.T_MAX)
    wv = w * ((idx >= 0) & (idx < S))
    idc = idx.clamp(0, S - 1)
    rows = (x[:, idc, :] * wv[None, :, :, None]).sum(2)
    out = (rows[:, :, idc] * wv[None, None, :, :]).sum(3)
    assert th.allclose(out, th.from_numpy(d[name + "_y"])[0], atol=1e-6)
    inv = rr.inverse_ranges(left, T, S)
    for r in (0, S // 3, S - 1):  # every output whose field of view holds r lies in [lo, hi], and none outside
        touching = [o for o in range(O) if left[o] <= r <= left[o] + T - 1]
        assert touching == list(range(int(inv[r, 0]), int(inv[r, 1]) + 1))


# ---------------------------------------------------------------------- use_augs (cgd/modules.py:12-24, 62), pinned on the reference
def _aug_cases():
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augs_golden.npz"))
    B, H, W, CS, CUTN = (int(v) for v in d["meta"])
    return d, B, H, W, CS, CUTN


def test_aug_draw_order_and_oracle_reproduce_the_reference_bit_exactly():
    """tests/golden/augs_golden.npz was written by the reference's MakeCutouts(use_augs=True) (torchvision transforms, CPU).  From
    the same seed, clip_guided_diffusion_b200/augs.py must consume the default generator identically (same final RNG state, same
    parameters and noise) and oracle.guidance.apply_augs must reproduce the reference's cutouts and input gradient bit for bit."""
    from clip_guided_diffusion_b200 import augs
    from oracle import guidance as og
    d, B, H, W, CS, CUTN = _aug_cases()
    for s in d["seeds"]:
        s = int(s)
        x, cot = th.from_numpy(d[f"x_{s}"]), th.from_numpy(d[f"cot_{s}"])
        mk = og.MakeCutouts(CS, CUTN, 1.0, use_augs=True)
        th.manual_seed(s)
        coords = mk._generate_coords(H, W, CUTN)
        noise = th.zeros(CUTN, 4, B, 3, min(H, W), min(H, W))
        prm = augs.draw_aug_params(coords, B, H, W, noise_device="cpu", noise_out=noise)
        assert th.equal(th.get_rng_state(), th.from_numpy(d[f"rng_{s}"]))
        assert coords == [tuple(int(v) for v in r) for r in d[f"coords_{s}"]]
        assert th.equal(prm, th.from_numpy(d[f"prm_{s}"])) and th.equal(noise, th.from_numpy(d[f"noise_{s}"]))
        xr = x.clone().requires_grad_()
        y = mk(xr, coords=coords, aug_params=prm, aug_noise=noise)
        (gx,) = th.autograd.grad((y * cot).sum(), xr)
        assert th.equal(y.detach(), th.from_numpy(d[f"y_{s}"])), s
        assert th.equal(gx, th.from_numpy(d[f"gx_{s}"])), s


def test_aug_ops_of_the_interpreter_match_the_reference_goldens():
    """the CUTOUTS_AUG_FWD / _BWD op contract (what the CUDA kernels implement), executed by tests/plan_interp.py"""
    from clip_guided_diffusion_b200.plan import Plan
    from tests.plan_interp import Interp
    d, B, H, W, CS, CUTN = _aug_cases()
    s = int(d["seeds"][0])
    P, kpad = CS, 3 * CS * CS
    plan = Plan()
    bx, bc, bp = plan.new(B * 3 * H * W, "f", "x"), plan.new(CUTN * 3, "i32", "coords"), plan.new(CUTN * 20, "f", "prm")
    bn = plan.new(CUTN * 4 * B * 3 * H * W, "f", "noise")
    bo, bd, bg = plan.new(CUTN * B * kpad, "h", "patches"), plan.new(CUTN * B * kpad, "h", "dpatches"), plan.new(B * 3 * H * W, "f", "dx")
    plan.emit("CUTOUTS_AUG_FWD", i=[B, H, W, CUTN, CS, P, kpad, min(H, W)], f=[0, 0, 0, 1, 1, 1], p=[(bx, 0), (bc, 0), (bo, 0), (bp, 0), (bn, 0)])
    plan.emit("FILL", i=[B * 3 * H * W], f=[0.0], p=[(bg, 0)])
    plan.emit("CUTOUTS_AUG_BWD", i=[B, H, W, CUTN, CS, P, kpad], f=[0, 0, 0, 1, 1, 1, 1.0], p=[(bd, 0), (bc, 0), (bg, 0), (bp, 0)])
    plan.finalize("cpu")
    plan.view(bx, (B, 3, H, W)).copy_(th.from_numpy(d[f"x_{s}"]) * 2 - 1)  # the op takes x_in in [-1, 1] and applies (x + 1) / 2
    plan.view(bc, (CUTN, 3)).copy_(th.from_numpy(d[f"coords_{s}"]))
    plan.view(bp, (CUTN, 20)).copy_(th.from_numpy(d[f"prm_{s}"]))
    plan.view(bn).copy_(th.from_numpy(d[f"noise_{s}"]).flatten())
    plan.view(bd, (CUTN * B, 1, kpad)).copy_(th.from_numpy(d[f"cot_{s}"]).reshape(CUTN * B, 1, kpad))
    Interp(plan).run(0, 3)
    y = plan.view(bo, (CUTN * B, 3, CS, CS)).float()
    ref = th.from_numpy(d[f"y_{s}"])
    assert float((y - ref).abs().max()) < 2e-3  # fp16 output
    gx = plan.view(bg, (B, 3, H, W))
    gref = th.from_numpy(d[f"gx_{s}"]) * 0.5  # d/dx_in of ((x_in + 1) / 2); fp16 cotangent
    assert float((gx - gref).abs().max() / gref.abs().max()) < 2e-3


def test_checkpoint_table_matches_the_reference_data_module():
    """`config_for` (product and oracle), the checkpoint file names and the CLIP tower list against the reference's own tables
    (data/diffusion_model_flags.py, cgd/clip_util.py:17-29; golden written by tests/golden/make_golden_flags.py), merged over
    guided-diffusion's `model_and_diffusion_defaults()` the way cgd/script_util.py:305-315 does."""
    import json
    import os
    from clip_guided_diffusion_b200 import cgd as pcgd
    from clip_guided_diffusion_b200 import unet as pu
    from clip_guided_diffusion_b200.rn import RN_CONFIGS
    from clip_guided_diffusion_b200.vit import VIT_CONFIGS
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_flags_golden.json")))
    defaults = dict(num_heads=4, num_head_channels=-1, use_new_attention_order=False, rescale_timesteps=False, noise_schedule="linear",
                    class_cond=False)  # [3P] guided_diffusion/script_util.py model_and_diffusion_defaults
    mult = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}  # channel_mult "" -> by image size
    assert len(gold["diffusion"]) == 6
    for key, entry in gold["diffusion"].items():
        cond, size = key.split("/")[0] == "cond", int(key.split("/")[1])
        fl = {**defaults, **entry["model_flags"]}
        assert fl["image_size"] == size and fl["class_cond"] == cond
        # what every kernel of the path assumes about the published checkpoints
        assert fl["learn_sigma"] and fl["resblock_updown"] and fl["use_scale_shift_norm"] and fl["use_fp16"] and fl["diffusion_steps"] == 1000
        assert pcgd.DIFFUSION_FILENAMES[(cond, size)] == entry["filename"]
        for cfg in (pu.config_for(size, cond), config_for(size, cond)):
            assert cfg.image_size == size and cfg.class_cond == cond
            assert cfg.model_channels == fl["num_channels"] and cfg.num_res_blocks == fl["num_res_blocks"]
            assert tuple(cfg.attention_resolutions) == tuple(int(r) for r in fl["attention_resolutions"].split(","))
            assert tuple(cfg.channel_mult) == mult[size]
            assert cfg.num_head_channels == fl["num_head_channels"] and (fl["num_head_channels"] != -1 or cfg.num_heads == fl["num_heads"])
            assert cfg.use_new_attention_order == fl["use_new_attention_order"]
            assert cfg.rescale_timesteps == fl["rescale_timesteps"]
        assert pu.config_for(size, cond).noise_schedule == fl["noise_schedule"]
    towers = {**VIT_CONFIGS, **RN_CONFIGS}
    assert set(gold["clip"]["names"]) <= set(towers), "every name of CLIP_MODEL_NAMES has a tower configuration"
    for name, fname in gold["clip"]["files"].items():
        assert pcgd.clip_checkpoint_filename(name) == fname
