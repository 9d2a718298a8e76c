"""fp32 PyTorch restatement of guided-diffusion's ADM ``UNetModel`` (oracle; test infrastructure).

The network itself is third-party (crowsonkb/guided-diffusion@fb47224,
``guided_diffusion/unet.py`` + ``nn.py``) and is absent from /root/reference; the
reference only constructs and calls it (``cgd/script_util.py:305-324``,
``cgd/cgd.py:250-262``).  This file restates the published architecture
 keeping the upstream ``state_dict`` key names
(Appendix A.5) so real checkpoints would load.  PARITY UNPINNED (see
``oracle/__init__.py``); structural pins: parameter counts in
``tests/test_oracle.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch as th
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
@dataclass
class UNetConfig:
    """Shape-defining flags (reference: ``data/diffusion_model_flags.py:1-120`` merged over
    guided-diffusion's ``model_and_diffusion_defaults()``, ``cgd/script_util.py:305-315``)."""

    image_size: int = 256
    model_channels: int = 256
    num_res_blocks: int = 2
    channel_mult: tuple = (1, 1, 2, 2, 4, 4)
    attention_resolutions: tuple = (32, 16, 8)  # spatial resolutions that carry attention
    num_heads: int = 4
    num_head_channels: int = 64  # -1 -> use num_heads
    class_cond: bool = True
    num_classes: int = 1000
    use_new_attention_order: bool = False
    in_channels: int = 3
    out_channels: int = 6  # learn_sigma=True in every flag set
    rescale_timesteps: bool = False
    noise_schedule: str = "linear"

    @property
    def attention_ds(self):
        return tuple(self.image_size // int(r) for r in self.attention_resolutions)


def default_channel_mult(image_size: int):
    # guided_diffusion/script_util.py create_model()
    return {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}[image_size]


def config_for(image_size: int, class_cond: bool = True) -> UNetConfig:
    """The four checkpoints of ``data/diffusion_model_flags.py`` (64/128/256/512)."""
    base = dict(image_size=image_size, class_cond=class_cond, channel_mult=default_channel_mult(image_size))
    if image_size == 64:
        return UNetConfig(model_channels=192, num_res_blocks=3, num_head_channels=64,
                          use_new_attention_order=True, noise_schedule="cosine", **base)
    if image_size == 128:
        return UNetConfig(model_channels=256, num_res_blocks=2, num_heads=4, num_head_channels=-1, **base)
    if image_size == 256:
        return UNetConfig(model_channels=256, num_res_blocks=2, num_head_channels=64, **base)
    if image_size == 512:
        return UNetConfig(model_channels=256, num_res_blocks=2, num_head_channels=64, rescale_timesteps=True, **base)
    raise ValueError(image_size)


def tiny_config(image_size=32, model_channels=64, channel_mult=(1, 2), num_res_blocks=1,
                attention_resolutions=(16,), class_cond=True, num_classes=10,
                use_new_attention_order=False) -> UNetConfig:
    """A small same-topology network for parity tests the oracle finishes in seconds."""
    return UNetConfig(image_size=image_size, model_channels=model_channels, num_res_blocks=num_res_blocks,
                      channel_mult=tuple(channel_mult), attention_resolutions=tuple(attention_resolutions),
                      num_head_channels=64, class_cond=class_cond, num_classes=num_classes,
                      use_new_attention_order=use_new_attention_order)


# --------------------------------------------------------------------------- pieces
def timestep_embedding(t: th.Tensor, dim: int, max_period: float = 10000.0) -> th.Tensor:
    """[cos | sin] sinusoid, frequencies exp(-ln(max_period) * i / half) (Appendix A.1)."""
    half = dim // 2
    freqs = th.exp(-math.log(max_period) * th.arange(half, dtype=th.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = th.cat([th.cos(args), th.sin(args)], dim=-1)
    if dim % 2:
        emb = th.cat([emb, th.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class EmbedSequential(nn.Sequential):
    """Sequential whose ResBlock children also receive the timestep embedding."""

    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, ResBlock) else layer(x)
        return x


class _Up(nn.Module):
    def forward(self, x):
        return F.interpolate(x, scale_factor=2, mode="nearest")


class _Down(nn.Module):
    def forward(self, x):
        return F.avg_pool2d(x, 2)


class ResBlock(nn.Module):
    """GN-SiLU-conv3x3, scale-shift GN, SiLU-conv3x3, skip (identity or 1x1); optional up/down
    resampling applied after the first GN+SiLU to both branches (resblock_updown=True)."""

    def __init__(self, channels, emb_channels, out_channels=None, up=False, down=False):
        super().__init__()
        out_channels = out_channels or channels
        self.channels, self.out_channels = channels, out_channels
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(), nn.Conv2d(channels, out_channels, 3, padding=1))
        self.updown = up or down
        if up:
            self.h_upd, self.x_upd = _Up(), _Up()
        elif down:
            self.h_upd, self.x_upd = _Down(), _Down()
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, out_channels), nn.SiLU(), nn.Dropout(p=0.0),
                                        nn.Conv2d(out_channels, out_channels, 3, padding=1))
        self.skip_connection = nn.Identity() if out_channels == channels else nn.Conv2d(channels, out_channels, 1)

    def forward(self, x, emb):
        if self.updown:
            h = self.in_layers[1](self.in_layers[0](x))
            h = self.h_upd(h)
            x = self.x_upd(x)
            h = self.in_layers[2](h)
        else:
            h = self.in_layers(x)
        e = self.emb_layers(emb).type(h.dtype)[:, :, None, None]
        scale, shift = th.chunk(e, 2, dim=1)
        h = self.out_layers[0](h) * (1 + scale) + shift
        h = self.out_layers[3](self.out_layers[2](self.out_layers[1](h)))
        return self.skip_connection(x) + h


class AttentionBlock(nn.Module):
    """GN -> qkv 1x1 -> multi-head softmax attention over H*W tokens -> proj 1x1 -> residual."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_new_attention_order=False):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads if num_head_channels == -1 else channels // num_head_channels
        self.new_order = use_new_attention_order
        self.norm = GroupNorm32(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = nn.Conv1d(channels, channels, 1)

    def forward(self, x):
        b, c = x.shape[:2]
        xf = x.reshape(b, c, -1)
        qkv = self.qkv(self.norm(xf))
        t = qkv.shape[-1]
        nh = self.num_heads
        d = c // nh
        if self.new_order:  # q,k,v = chunk(3, dim=1), then heads
            q, k, v = qkv.chunk(3, dim=1)
            q, k, v = (z.reshape(b * nh, d, t) for z in (q, k, v))
        else:  # legacy: per-head [q|k|v] interleave
            q, k, v = qkv.reshape(b * nh, 3 * d, t).split(d, dim=1)
        s = 1.0 / math.sqrt(math.sqrt(d))
        w = th.einsum("bct,bcs->bts", q * s, k * s)
        w = th.softmax(w.float(), dim=-1).type(w.dtype)
        a = th.einsum("bts,bcs->bct", w, v).reshape(b, c, t)
        return (xf + self.proj_out(a)).reshape(x.shape)


# --------------------------------------------------------------------------- model
class UNetModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        mc = cfg.model_channels
        ted = mc * 4
        self.model_channels = mc
        self.num_classes = cfg.num_classes if cfg.class_cond else None
        self.dtype = th.float32
        self.time_embed = nn.Sequential(nn.Linear(mc, ted), nn.SiLU(), nn.Linear(ted, ted))
        if self.num_classes is not None:
            self.label_emb = nn.Embedding(self.num_classes, ted)

        attn_kw = dict(num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                       use_new_attention_order=cfg.use_new_attention_order)
        ch = int(cfg.channel_mult[0] * mc)
        self.input_blocks = nn.ModuleList([EmbedSequential(nn.Conv2d(cfg.in_channels, ch, 3, padding=1))])
        skip_chs = [ch]
        ds = 1
        for level, mult in enumerate(cfg.channel_mult):
            for _ in range(cfg.num_res_blocks):
                layers = [ResBlock(ch, ted, int(mult * mc))]
                ch = int(mult * mc)
                if ds in cfg.attention_ds:
                    layers.append(AttentionBlock(ch, **attn_kw))
                self.input_blocks.append(EmbedSequential(*layers))
                skip_chs.append(ch)
            if level != len(cfg.channel_mult) - 1:
                self.input_blocks.append(EmbedSequential(ResBlock(ch, ted, ch, down=True)))
                skip_chs.append(ch)
                ds *= 2
        self.middle_block = EmbedSequential(ResBlock(ch, ted), AttentionBlock(ch, **attn_kw), ResBlock(ch, ted))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
            for i in range(cfg.num_res_blocks + 1):
                ich = skip_chs.pop()
                layers = [ResBlock(ch + ich, ted, int(mc * mult))]
                ch = int(mc * mult)
                if ds in cfg.attention_ds:
                    layers.append(AttentionBlock(ch, **attn_kw))
                if level and i == cfg.num_res_blocks:
                    layers.append(ResBlock(ch, ted, ch, up=True))
                    ds //= 2
                self.output_blocks.append(EmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(ch, cfg.out_channels, 3, padding=1))

    def forward(self, x, timesteps, y=None):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        if self.num_classes is not None:
            assert y is not None and y.shape == (x.shape[0],)
            emb = emb + self.label_emb(y)
        hs = []
        h = x.type(self.dtype)
        for m in self.input_blocks:
            h = m(h, emb)
            hs.append(h)
        h = self.middle_block(h, emb)
        for m in self.output_blocks:
            h = m(th.cat([h, hs.pop()], dim=1), emb)
        return self.out(h.type(x.dtype))


def seeded_init_(model: nn.Module, seed: int = 1234, gain: float = 1.0) -> nn.Module:
    """Deterministic fp16-range-safe weights in upstream layout: conv/linear
    N(0, gain/fan_in) -- including upstream's zero-initialised layers, which would otherwise make
    the network trivially zero -- biases N(0, 0.02), norm gains 1+N(0,0.1), embeddings N(0,0.5)."""
    g = th.Generator().manual_seed(seed)
    with th.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2 and "label_emb" not in name and "positional_embedding" not in name and name != "proj" and not name.endswith(".proj"):
                fan_in = p[0].numel()
                p.copy_(th.randn(p.shape, generator=g) * math.sqrt(gain / fan_in))
            elif "label_emb" in name:
                p.copy_(th.randn(p.shape, generator=g) * 0.5)
            elif p.dim() >= 2:
                p.copy_(th.randn(p.shape, generator=g) * (p.shape[0] ** -0.5))
            elif name.endswith("bias"):
                p.copy_(th.randn(p.shape, generator=g) * 0.02)
            elif name.endswith("weight"):  # norm gains
                p.copy_(1.0 + 0.1 * th.randn(p.shape, generator=g))
            else:  # class_embedding etc.
                p.copy_(th.randn(p.shape, generator=g) * 0.1)
    return model
