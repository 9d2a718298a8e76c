"""fp32 restatement of CLIP's visual ``ModifiedResNet`` (RN50 family; oracle, test infrastructure).

Third-party (clip-anytorch 2.6.0 ``clip/model.py``: ``Bottleneck``, ``AttentionPool2d``, ``ModifiedResNet``), absent from
/root/reference; the reference offers these towers through ``clip.load(clip_model_name)`` (cgd/clip_util.py:17, 59-64) and calls
``clip_model.encode_image`` at cgd/cgd.py:194.  Restated from the published architecture with upstream state_dict keys
(``visual.conv1 .. visual.layer4.*, visual.attnpool.*``).  PARITY UNPINNED (no package, no checkpoint here); structural pin: RN50's
visual tower has 38.3 M parameters.

    stem: conv3x3 s2 (3 -> w/2), conv3x3 (w/2 -> w/2), conv3x3 (w/2 -> w), each + BatchNorm + ReLU; AvgPool2d(2)
    Bottleneck(inplanes, planes, stride): conv1x1 -> BN -> ReLU -> conv3x3 -> BN -> ReLU -> AvgPool2d(stride) -> conv1x1 (4 planes) -> BN,
        identity through AvgPool2d(stride) -> conv1x1 -> BN when the shape changes; ReLU(out + identity)
    layers (3, 4, 6, 3) at widths (w, 2w, 4w, 8w), strides (1, 2, 2, 2)
    AttentionPool2d: tokens = [mean, x_1 .. x_HW] + positional_embedding; multi-head attention with query = the mean token only;
        c_proj -> output_dim
BatchNorm is in eval mode (the reference calls ``clip_model.eval()``, cgd/clip_util.py:64).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import torch as th
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class RNConfig:
    layers: tuple = (3, 4, 6, 3)
    output_dim: int = 1024
    input_resolution: int = 224
    width: int = 64

    @property
    def heads(self):
        return self.width * 32 // 64

    @property
    def embed_dim(self):
        return self.width * 32


RN_CONFIGS = {"RN50": RNConfig((3, 4, 6, 3), 1024, 224, 64), "RN101": RNConfig((3, 4, 23, 3), 512, 224, 64)}


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if stride > 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(OrderedDict([("-1", nn.AvgPool2d(stride)), ("0", nn.Conv2d(inplanes, planes * 4, 1, bias=False)),
                                                         ("1", nn.BatchNorm2d(planes * 4))]))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(self.avgpool(out)))
        identity = x if self.downsample is None else self.downsample(x)
        return F.relu(out + identity)


class AttentionPool2d(nn.Module):
    def __init__(self, spacial_dim, embed_dim, num_heads, output_dim):
        super().__init__()
        self.positional_embedding = nn.Parameter(th.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim)
        self.num_heads = num_heads

    def forward(self, x):
        x = x.flatten(start_dim=2).permute(2, 0, 1)  # NCHW -> (HW)NC
        x = th.cat([x.mean(dim=0, keepdim=True), x], dim=0)
        x = x + self.positional_embedding[:, None, :].to(x.dtype)
        T, n, C = x.shape
        d = C // self.num_heads
        q = self.q_proj(x[:1]).view(1, n * self.num_heads, d).transpose(0, 1) * d ** -0.5
        k = self.k_proj(x).view(T, n * self.num_heads, d).transpose(0, 1)
        v = self.v_proj(x).view(T, n * self.num_heads, d).transpose(0, 1)
        a = th.softmax(q @ k.transpose(1, 2), dim=-1) @ v  # [n*heads, 1, d]
        return self.c_proj(a.transpose(0, 1).reshape(1, n, C)).squeeze(0)


class ModifiedResNet(nn.Module):
    def __init__(self, cfg: RNConfig):
        super().__init__()
        self.cfg = cfg
        self.input_resolution, self.output_dim = cfg.input_resolution, cfg.output_dim
        w = cfg.width
        self.conv1 = nn.Conv2d(3, w // 2, 3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(w // 2)
        self.conv2 = nn.Conv2d(w // 2, w // 2, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(w // 2)
        self.conv3 = nn.Conv2d(w // 2, w, 3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(w)
        self.avgpool = nn.AvgPool2d(2)
        self._inplanes = w
        self.layer1 = self._make_layer(w, cfg.layers[0])
        self.layer2 = self._make_layer(w * 2, cfg.layers[1], stride=2)
        self.layer3 = self._make_layer(w * 4, cfg.layers[2], stride=2)
        self.layer4 = self._make_layer(w * 8, cfg.layers[3], stride=2)
        self.attnpool = AttentionPool2d(cfg.input_resolution // 32, cfg.embed_dim, cfg.heads, cfg.output_dim)

    def _make_layer(self, planes, blocks, stride=1):
        layers = [Bottleneck(self._inplanes, planes, stride)]
        self._inplanes = planes * Bottleneck.expansion
        layers += [Bottleneck(self._inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = x.type(self.conv1.weight.dtype)
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = F.relu(bn(conv(x)))
        x = self.avgpool(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.attnpool(x)


class CLIPVisualRN(nn.Module):
    """``encode_image`` + ``visual`` of a CLIP model with a ModifiedResNet tower"""

    def __init__(self, cfg: RNConfig):
        super().__init__()
        self.visual = ModifiedResNet(cfg)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))


def seeded_init_rn_(model: nn.Module, seed: int = 77) -> nn.Module:
    """He-initialised convs / linears (the activations stay O(1) through ~50 ReLU layers), non-trivial BatchNorm statistics"""
    g = th.Generator().manual_seed(seed)
    with th.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and "positional_embedding" not in name:
                # the last conv of every block is damped: 16 residual additions would otherwise double the variance 16 times
                gain = 0.25 if name.endswith("conv3.weight") else 1.0
                p.copy_(th.randn(p.shape, generator=g) * gain * (2.0 / p[0].numel()) ** 0.5)
            elif "positional_embedding" in name:
                p.copy_(th.randn(p.shape, generator=g) * p.shape[1] ** -0.5)
            elif name.endswith("bias"):
                p.copy_(th.randn(p.shape, generator=g) * 0.05)
            else:  # BatchNorm gains
                p.copy_(1.0 + 0.1 * th.randn(p.shape, generator=g))
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(th.randn(b.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                b.copy_(1.0 + 0.2 * th.rand(b.shape, generator=g))
    return model.eval()
