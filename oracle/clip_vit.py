"""fp32 restatement of CLIP's visual ``VisionTransformer`` (oracle; test infrastructure).

Third-party (clip-anytorch 2.6.0 ``clip/model.py``), absent from /root/reference; the reference
calls ``clip_model.encode_image`` at ``cgd/cgd.py:194`` and reads ``visual.input_resolution`` at
``cgd/clip_util.py:61,66``.  Restated from SURVEY.md Appendix A.3 with upstream state_dict keys
(``visual.*``).  PARITY UNPINNED (no upstream vectors); structural pins = parameter counts (87.8 / 86.2 / 304.0 M);
CROSS-CHECKED against the independent HuggingFace ``transformers`` CLIP vision tower with re-keyed seeded weights
(tests/test_oracle_crosscheck.py: forward and input gradient agree to 2e-5 for a tiny tower, ViT-B/32 and ViT-B/16).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import torch as th
import torch.nn as nn


@dataclass
class ViTConfig:
    input_resolution: int = 224
    patch_size: int = 32
    width: int = 768
    layers: int = 12
    output_dim: int = 512

    @property
    def heads(self):
        return self.width // 64

    @property
    def tokens(self):
        return (self.input_resolution // self.patch_size) ** 2 + 1


VIT_CONFIGS = {
    "ViT-B/32": ViTConfig(224, 32, 768, 12, 512),
    "ViT-B/16": ViTConfig(224, 16, 768, 12, 512),
    "ViT-L/14": ViTConfig(224, 14, 1024, 24, 768),
    "ViT-L/14@336px": ViTConfig(336, 14, 1024, 24, 768),  # in CLIP_MODEL_URLS (cgd/clip_util.py:28): 24 x 24 patches + class token = 577 tokens
}


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * th.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)

    def forward(self, x):  # x: [T, n, w]
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, cfg: ViTConfig):
        super().__init__()
        self.cfg = cfg
        self.input_resolution = cfg.input_resolution
        self.output_dim = cfg.output_dim
        w = cfg.width
        self.conv1 = nn.Conv2d(3, w, cfg.patch_size, stride=cfg.patch_size, bias=False)
        s = w ** -0.5
        self.class_embedding = nn.Parameter(s * th.randn(w))
        self.positional_embedding = nn.Parameter(s * th.randn(cfg.tokens, w))
        self.ln_pre = LayerNorm(w)
        self.transformer = Transformer(w, cfg.layers, cfg.heads)
        self.ln_post = LayerNorm(w)
        self.proj = nn.Parameter(s * th.randn(w, cfg.output_dim))

    def forward(self, x):
        x = self.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # [n, g*g, w]
        cls = self.class_embedding.to(x.dtype) + th.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = th.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


class CLIPVisualOnly(nn.Module):
    """The slice of ``clip.model.CLIP`` the hot path touches: ``encode_image`` and ``visual``."""

    def __init__(self, cfg: ViTConfig):
        super().__init__()
        self.visual = VisionTransformer(cfg)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))
