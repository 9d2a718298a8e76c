"""CPU: the CLIP ModifiedResNet tower (rn.py) as an op list, interpreted, against the fp32 oracle restatement (oracle/clip_rn.py):
folded BatchNorm, the stride-2 stem conv as a stride-1 conv over the space-to-depth cutouts, bottlenecks with average-pool
downsampling, AttentionPool2d; forward embeddings and the input gradient."""
import pytest
import torch as th

from clip_guided_diffusion_b200 import rn as prn
from clip_guided_diffusion_b200 import weights as pw
from oracle import clip_rn as orn
from tests.plan_interp import Interp


def test_rn50_structure():
    sh = pw.rn_param_shapes(prn.RN_CONFIGS["RN50"])
    m = orn.ModifiedResNet(orn.RN_CONFIGS["RN50"])
    ref = {"visual." + k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    assert sh == ref
    n_params = sum(th.Size(s).numel() for k, s in sh.items() if "running_" not in k)
    assert n_params == 38316896  # the published RN50 visual tower
    assert prn.rn_config_from_state_dict({k: th.empty(s) for k, s in sh.items()}) == prn.RN_CONFIGS["RN50"]


def test_stem_space_to_depth_weight():
    """the stride-2 3x3 conv equals a stride-1 3x3 conv over the 2x2 space-to-depth image with the remapped kernel"""
    import torch.nn.functional as F
    g = th.Generator().manual_seed(0)
    w = th.randn(5, 3, 3, 3, generator=g)
    x = th.randn(2, 3, 16, 16, generator=g)
    ref = F.conv2d(x, w, stride=2, padding=1)
    s2d = x.view(2, 3, 8, 2, 8, 2).permute(0, 1, 3, 5, 2, 4).reshape(2, 12, 8, 8)  # channel = (c, ky, kx)
    got = F.conv2d(s2d, prn.stem_s2d_weight(w), padding=1)
    assert th.allclose(got, ref, atol=1e-5)


def test_published_wide_towers_structure():
    """RN50x4 / RN50x16 (cgd/clip_util.py:17): published depth, width, resolution and embedding size; widths 80 / 96 are not multiples
    of the 64-channel K slice (rn._conv_bn zero-pads them)"""
    for name, (n_params, tokens, embed) in {"RN50x4": (None, 82, 2560), "RN50x16": (None, 145, 3072)}.items():
        cfg = prn.RN_CONFIGS[name]
        sh = pw.rn_param_shapes(cfg)
        m = orn.ModifiedResNet(orn.RNConfig(tuple(cfg.layers), cfg.output_dim, cfg.input_resolution, cfg.width))
        ref = {"visual." + k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
        assert sh == ref
        assert cfg.tokens == tokens and cfg.embed_dim == embed and cfg.embed_dim // cfg.heads == 64
        assert prn.rn_config_from_state_dict({k: th.empty(s) for k, s in sh.items()}) == cfg


@pytest.mark.parametrize("n,width", [(2, 64), (2, 80), (1, 48)])
def test_rn_tower_matches_oracle(n, width):
    cfg = prn.RNConfig(layers=(1, 2, 1, 1), output_dim=128, input_resolution=64, width=width)
    ocfg = orn.RNConfig(layers=(1, 2, 1, 1), output_dim=128, input_resolution=64, width=width)
    sd = pw.seeded_rn_state_dict(cfg, seed=3)
    oracle = orn.ModifiedResNet(ocfg).eval()
    missing = oracle.load_state_dict({k[len("visual."):]: v for k, v in sd.items()}, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys) and not missing.unexpected_keys
    tower = prn.RNB200(cfg, sd, n_images=n, device="cpu")
    it = Interp(tower.plan)
    g = th.Generator().manual_seed(1)
    img = th.randn(n, 3, 64, 64, generator=g)
    tower.plan.view(tower.patches, (n, cfg.grid ** 2, cfg.kpad)).copy_(Interp._patchify(img, 2, cfg.kpad))
    it.run_range("vit_fwd", "vit_bwd")
    got = tower.plan.view(tower.embeds, (n, cfg.output_dim)).clone()
    xi = img.clone().requires_grad_()
    ref = oracle(xi)
    assert float((got - ref.detach()).norm() / ref.detach().norm()) < 2e-2
    d_emb = th.randn(n, cfg.output_dim, generator=g)
    tower.plan.view(tower.d_embeds, (n, cfg.output_dim)).copy_(d_emb)
    it.run_range("vit_bwd", "vit_end")
    dp = tower.plan.view(tower.d_patches, (n, cfg.grid ** 2, cfg.kpad)).float()
    d_img = Interp._unpatchify(dp, 2, 64)
    (g_ref,) = th.autograd.grad((ref * d_emb).sum(), xi)
    cos = float(th.nn.functional.cosine_similarity(d_img.flatten(), g_ref.flatten(), dim=0))
    assert cos > 0.995 and float((d_img - g_ref).norm() / g_ref.norm()) < 5e-2, cos


def test_oracle_attention_pool_equals_torch_mha():
    """the oracle's AttentionPool2d restatement against the call the published model makes (F.multi_head_attention_forward with separate
    projection weights, query = the mean token)"""
    import torch.nn.functional as F
    th.manual_seed(0)
    ap = orn.AttentionPool2d(7, 2048, 32, 1024)
    f = th.randn(2, 2048, 7, 7)
    x = f.flatten(start_dim=2).permute(2, 0, 1)
    x = th.cat([x.mean(dim=0, keepdim=True), x], dim=0) + ap.positional_embedding[:, None, :]
    with th.no_grad():
        ref, _ = F.multi_head_attention_forward(
            query=x[:1], key=x, value=x, embed_dim_to_check=x.shape[-1], num_heads=ap.num_heads, q_proj_weight=ap.q_proj.weight,
            k_proj_weight=ap.k_proj.weight, v_proj_weight=ap.v_proj.weight, in_proj_weight=None,
            in_proj_bias=th.cat([ap.q_proj.bias, ap.k_proj.bias, ap.v_proj.bias]), bias_k=None, bias_v=None, add_zero_attn=False, dropout_p=0,
            out_proj_weight=ap.c_proj.weight, out_proj_bias=ap.c_proj.bias, use_separate_proj_weight=True, training=False, need_weights=False)
        got = ap(f)
    assert th.allclose(got, ref.squeeze(0), atol=1e-5)
