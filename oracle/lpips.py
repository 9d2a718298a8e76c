"""fp32 restatement of ``lpips.LPIPS(net='vgg')`` v0.1 in eval mode (oracle; test infrastructure).

The reference builds it lazily (cgd/cgd.py:147-148) and calls ``lpips_vgg(x_in, init_tensor).sum() * init_scale``
(cgd/cgd.py:220-224) with inputs in [-1, 1].  The ``lpips`` package (pin 0.1.4, uv.lock:609-610) and torchvision's VGG16
weights are NOT in /root/reference and cannot be installed here: this follows the published algorithm and the
upstream state_dict key layout -- PARITY UNPINNED (14.7 M VGG16 feature parameters + 1472 lin weights); the VGG16 trunk and its
five tap positions are CROSS-CHECKED against torchvision.models.vgg16().features (the network lpips wraps) with the same seeded weights
(tests/test_oracle_crosscheck.py).

    ScalingLayer: (x - shift) / scale, shift = [-.030, -.088, -.188], scale = [.458, .448, .450]
    VGG16 features, taps after relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 (64, 128, 256, 512, 512 channels)
    per tap: x / (||x||_2 over channels + 1e-10), squared difference, 1x1 conv with non-negative weights (no bias),
    spatial mean; sum over taps -> [N, 1, 1, 1]
"""
from __future__ import annotations

import torch as th
import torch.nn as nn
import torch.nn.functional as F

SHIFT = (-0.030, -0.088, -0.188)
SCALE = (0.458, 0.448, 0.450)
# torchvision vgg16.features indices of the convs of each LPIPS slice (slice k ends at relu after its last conv)
SLICES = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21), (24, 26, 28))
CHANNELS = (64, 128, 256, 512, 512)


def param_shapes() -> dict:
    """upstream key -> shape (``net.sliceK.<features index>.*`` as lpips.pretrained_networks.vgg16 registers them,
    ``linK.model.1.weight`` from lpips.NetLinLayer)"""
    sh, cin = {}, 3
    for k, idxs in enumerate(SLICES):
        for i in idxs:
            sh[f"net.slice{k + 1}.{i}.weight"] = (CHANNELS[k], cin, 3, 3)
            sh[f"net.slice{k + 1}.{i}.bias"] = (CHANNELS[k],)
            cin = CHANNELS[k]
        sh[f"lin{k}.model.1.weight"] = (1, CHANNELS[k], 1, 1)
    return sh


def seeded_state_dict(seed: int = 77) -> dict:
    """He-initialised convs (activations stay O(1) through 13 ReLU layers), non-negative lin weights like the trained ones"""
    g = th.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes().items():
        if name.startswith("lin"):
            sd[name] = th.rand(shape, generator=g) * 0.2
        elif name.endswith("weight"):
            sd[name] = th.randn(shape, generator=g) * (2.0 / (shape[1] * 9)) ** 0.5
        else:
            sd[name] = th.randn(shape, generator=g) * 0.05
    return sd


def normalize_tensor(x, eps=1e-10):
    return x / (th.sqrt(th.sum(x ** 2, dim=1, keepdim=True)) + eps)


class LPIPSVgg(nn.Module):
    def __init__(self, state_dict: dict):
        super().__init__()
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.register_buffer("shift", th.tensor(SHIFT).view(1, 3, 1, 1))
        self.register_buffer("scale", th.tensor(SCALE).view(1, 3, 1, 1))

    def features(self, x):
        h = (x - self.shift) / self.scale
        taps = []
        for k, idxs in enumerate(SLICES):
            if k > 0:
                h = F.max_pool2d(h, 2, 2)
            for i in idxs:
                h = F.relu(F.conv2d(h, self.sd[f"net.slice{k + 1}.{i}.weight"], self.sd[f"net.slice{k + 1}.{i}.bias"], padding=1))
            taps.append(h)
        return taps

    def forward(self, in0, in1):
        f0, f1 = self.features(in0), self.features(in1)
        val = 0
        for k in range(5):
            d = (normalize_tensor(f0[k]) - normalize_tensor(f1[k])) ** 2
            val = val + F.conv2d(d, self.sd[f"lin{k}.model.1.weight"]).mean([2, 3], keepdim=True)
        return val
