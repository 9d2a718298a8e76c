"""CPU: the third-party restatements of `oracle/` checked against INDEPENDENT implementations that are installed in this image.

The packages the reference imports for this arithmetic (clip-anytorch 2.6.0, lpips 0.1.4, guided-diffusion @ fb47224) are absent
and cannot be installed (no network), so `oracle/clip_vit.py` and `oracle/lpips.py` were written from the published algorithm
(SURVEY.md A.3 / A.4).  What IS installed: HuggingFace `transformers` (its own CLIP vision tower, written independently of
OpenAI's `clip/model.py` but loading the same checkpoints) and `torchvision.models.vgg16` (the very network `lpips` wraps).
Seeded random weights in the upstream key layout are re-keyed into those models; outputs AND input gradients must agree to fp32
round-off.  This pins the op order, the attention scaling, QuickGELU, pre/post LayerNorm placement, class-token handling, the
projection, and the VGG slice boundaries -- everything a transcription error could get wrong -- without any checkpoint.
"""
import pytest
import torch as th

from clip_guided_diffusion_b200 import weights as pw
from clip_guided_diffusion_b200.vit import VIT_CONFIGS as P_VIT_CONFIGS
from clip_guided_diffusion_b200.vit import ViTConfig as PViTConfig
from oracle import lpips as ol
from oracle.clip_vit import CLIPVisualOnly, ViTConfig


def _hf_vision_from_upstream(sd: dict, cfg: ViTConfig):
    """clip-anytorch `visual.*` keys -> transformers CLIPVisionModelWithProjection (in_proj split into q/k/v, proj transposed)"""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    w = cfg.width
    hf_cfg = CLIPVisionConfig(hidden_size=w, intermediate_size=4 * w, projection_dim=cfg.output_dim, num_hidden_layers=cfg.layers,
                              num_attention_heads=cfg.heads, image_size=cfg.input_resolution, patch_size=cfg.patch_size,
                              hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0)
    hf_cfg._attn_implementation = "eager"
    m = CLIPVisionModelWithProjection(hf_cfg).eval()
    new = {"vision_model.embeddings.class_embedding": sd["visual.class_embedding"],
           "vision_model.embeddings.patch_embedding.weight": sd["visual.conv1.weight"],
           "vision_model.embeddings.position_embedding.weight": sd["visual.positional_embedding"],
           "vision_model.pre_layrnorm.weight": sd["visual.ln_pre.weight"], "vision_model.pre_layrnorm.bias": sd["visual.ln_pre.bias"],
           "vision_model.post_layernorm.weight": sd["visual.ln_post.weight"], "vision_model.post_layernorm.bias": sd["visual.ln_post.bias"],
           "visual_projection.weight": sd["visual.proj"].t().contiguous()}
    for i in range(cfg.layers):
        s, d = f"visual.transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[s + "attn.in_proj_weight"].chunk(3, 0)
        bq, bk, bv = sd[s + "attn.in_proj_bias"].chunk(3, 0)
        for n, (ww, bb) in dict(q_proj=(wq, bq), k_proj=(wk, bk), v_proj=(wv, bv)).items():
            new[d + f"self_attn.{n}.weight"], new[d + f"self_attn.{n}.bias"] = ww, bb
        for a, b in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"),
                     ("mlp.c_proj", "mlp.fc2")):
            new[d + b + ".weight"], new[d + b + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]
    missing, unexpected = m.load_state_dict(new, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return m


def _fwd_and_grad(fn, x, seed_out):
    x = x.clone().requires_grad_()
    y = fn(x)
    (g,) = th.autograd.grad((y * seed_out).sum(), x)
    return y.detach(), g


@pytest.mark.parametrize("name", ["tiny", "ViT-B/32", "ViT-B/16"])
def test_oracle_vit_matches_transformers_clip_vision(name):
    pytest.importorskip("transformers")
    if name == "tiny":
        cfg, pcfg, n = ViTConfig(64, 16, 128, 2, 64), PViTConfig(64, 16, 128, 2, 64), 3
    else:
        p = P_VIT_CONFIGS[name]
        cfg, pcfg, n = ViTConfig(p.input_resolution, p.patch_size, p.width, p.layers, p.output_dim), p, 2
    sd = pw.seeded_state_dict(pw.vit_param_shapes(pcfg), 1235)  # the very state_dicts the GPU parity tests load
    ours = CLIPVisualOnly(cfg).eval()
    ours.load_state_dict(sd)
    hf = _hf_vision_from_upstream(sd, cfg)
    g = th.Generator().manual_seed(1)
    x = th.randn(n, 3, cfg.input_resolution, cfg.input_resolution, generator=g)
    seed = th.randn(n, cfg.output_dim, generator=g)
    y0, g0 = _fwd_and_grad(ours.encode_image, x, seed)
    y1, g1 = _fwd_and_grad(lambda t: hf(pixel_values=t).image_embeds, x, seed)
    assert y0.shape == y1.shape == (n, cfg.output_dim)
    ey = float((y0 - y1).abs().max() / y1.abs().max())
    eg = float((g0 - g1).abs().max() / g1.abs().max())
    assert ey < 2e-5 and eg < 2e-5, (ey, eg)


def test_oracle_lpips_trunk_matches_torchvision_vgg16():
    """`lpips.pretrained_networks.vgg16` slices torchvision's vgg16.features at [0:4], [4:9], [9:16], [16:23], [23:30]; the oracle's
    `net.sliceK.<idx>` keys carry torchvision's own layer indices, so the re-keying is the identity on the index."""
    tv = pytest.importorskip("torchvision")
    sd = ol.seeded_state_dict()
    net = tv.models.vgg16(weights=None).features.eval()
    new = {}
    for k, v in sd.items():
        if k.startswith("net.slice"):
            _, _, idx, kind = k.split(".")
            new[f"{idx}.{kind}"] = v
    net.load_state_dict(new, strict=True)  # strict: the 13 convs of VGG16 and nothing else
    taps_at = (3, 8, 15, 22, 29)  # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
    g = th.Generator().manual_seed(2)
    x = (th.rand(2, 3, 64, 64, generator=g) * 2 - 1).requires_grad_()
    model = ol.LPIPSVgg(sd)
    ours = model.features(x)
    h = (x - model.shift) / model.scale
    theirs = []
    for i, layer in enumerate(net):
        h = layer(h)
        if i in taps_at:
            theirs.append(h)
    assert len(ours) == len(theirs) == 5
    for a, b in zip(ours, theirs):
        assert a.shape == b.shape and float((a - b).abs().max() / b.abs().max()) < 1e-5
    # the distance on top of the taps, restated independently (channel-unit-normalise, squared difference, lin, spatial mean)
    y = (th.rand(1, 3, 64, 64, generator=g) * 2 - 1)
    d = model(x, y)
    fx, fy = ours, model.features(y)
    want = 0
    for k in range(5):
        nx = fx[k] / (fx[k].pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
        ny = fy[k] / (fy[k].pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
        want = want + ((nx - ny).pow(2) * sd[f"lin{k}.model.1.weight"]).sum(1, keepdim=True).mean([2, 3], keepdim=True)
    assert d.shape == (2, 1, 1, 1) and float((d - want).abs().max() / want.abs().max()) < 1e-5
    assert tuple(model.shift.flatten().tolist()) == pytest.approx((-0.030, -0.088, -0.188))
    assert tuple(model.scale.flatten().tolist()) == pytest.approx((0.458, 0.448, 0.450))


def test_oracle_rn_bottleneck_stages_match_torchvision():
    """`oracle/clip_rn.py` Bottleneck vs torchvision's: CLIP's ModifiedResNet differs from the torchvision ResNet only in the stem, the
    anti-aliasing average pools (stride > 1) and the attention pool.  Its stride-1 stage is torchvision's `resnet50().layer1` key for
    key (the "-1" pool of `downsample` holds no parameters), so the same state_dict must give the same output and input gradient; a
    strided bottleneck equals a stride-1 torchvision block evaluated in two halves with the pool in between."""
    import torchvision
    from oracle.clip_rn import Bottleneck, ModifiedResNet, RNConfig
    g = th.Generator().manual_seed(11)
    o = ModifiedResNet(RNConfig(layers=(3, 1, 1, 1), output_dim=64, input_resolution=32, width=64)).eval()
    sd = {}
    for k, v in o.layer1.state_dict().items():  # random weights AND random BatchNorm running statistics
        sd[k] = v if "num_batches" in k else (th.rand(v.shape, generator=g) + 0.5 if "running_var" in k else th.randn(v.shape, generator=g) * 0.2)
    o.layer1.load_state_dict(sd)
    tv = torchvision.models.resnet50(weights=None).layer1.eval()
    tv.load_state_dict(sd)  # identical keys
    x = th.randn(2, 64, 8, 8, generator=g)
    seed = th.randn(2, 256, 8, 8, generator=g)
    a, ga = _fwd_and_grad(o.layer1, x, seed)
    b, gb = _fwd_and_grad(tv, x, seed)
    assert float((a - b).abs().max() / b.abs().max()) < 1e-5 and float((ga - gb).abs().max() / gb.abs().max()) < 1e-5

    blk = Bottleneck(256, 128, stride=2).eval()  # first block of layer2
    bsd = {}
    for k, v in blk.state_dict().items():
        bsd[k] = v if "num_batches" in k else (th.rand(v.shape, generator=g) + 0.5 if "running_var" in k else th.randn(v.shape, generator=g) * 0.2)
    blk.load_state_dict(bsd)
    tvb = torchvision.models.resnet.Bottleneck(256, 128, stride=1, downsample=th.nn.Sequential(
        th.nn.Conv2d(256, 512, 1, bias=False), th.nn.BatchNorm2d(512))).eval()
    tvb.load_state_dict(bsd)
    pool = th.nn.AvgPool2d(2)

    def tv_strided(z):  # torchvision's layers, CLIP's pool placement: after the 3x3's ReLU on the main branch, before the 1x1 on the skip
        h = tvb.relu(tvb.bn2(tvb.conv2(tvb.relu(tvb.bn1(tvb.conv1(z))))))
        return tvb.relu(tvb.bn3(tvb.conv3(pool(h))) + tvb.downsample(pool(z)))

    x = th.randn(2, 256, 8, 8, generator=g)
    seed = th.randn(2, 512, 4, 4, generator=g)
    a, ga = _fwd_and_grad(blk, x, seed)
    b, gb = _fwd_and_grad(tv_strided, x, seed)
    assert float((a - b).abs().max() / b.abs().max()) < 1e-5 and float((ga - gb).abs().max() / gb.abs().max()) < 1e-5
