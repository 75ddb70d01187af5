"""fp32 CPU restatement of the third-party model under the hot path -- TEST INFRASTRUCTURE ONLY.

The reference drives ``diffusers==0.8.0`` (requirements.yaml:174, assertion text at
ptp_utils.py:573), which is not vendored in /root/reference and not installed here.  This file
restates the published architecture of its ``UNet2DConditionModel`` / ``AutoencoderKL`` encoder /
``DDIMScheduler`` for the Stable Diffusion 1.x config, with the
diffusers-0.8.0 class and attribute names the reference's monkey patch relies on:

  * a top-level child literally called ``up_blocks`` (ptp_utils.py:565-568),
  * attention modules whose class is literally ``CrossAttention`` with ``to_q/to_k/to_v`` (no bias),
    ``to_out`` (ModuleList[Linear, Dropout]), ``heads``, ``scale``, ``reshape_heads_to_batch_dim``,
    ``reshape_batch_dim_to_heads`` and ``forward(hidden_states, context=None, mask=None)``
    (ptp_utils.py:474-491,540,556),
  * ``unet(x, t, ctx)["sample"]`` (ptp_utils.py:227-229), ``vae.encode(img)["latent_dist"].mean``
    (ptp_utils.py:299-302), ``scheduler.timesteps`` / ``scheduler.add_noise`` (ptp_utils.py:221-223).

Parameter names follow the diffusers state-dict keys so real SD1.5 weights would load.
Parity vs real diffusers: UNPINNED (cannot be checked offline).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: int = 8  # diffusers 0.8.0 SD1.x: this is the NUMBER of heads
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    # which down/up stages carry transformers (SD1.x: first three down, last three up)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)

    @staticmethod
    def sd15() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def tiny() -> "UNetConfig":
        """Small same-topology config used for committed golden fixtures."""
        return UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=48,
                          attention_head_dim=4, norm_num_groups=8)


@dataclass
class VAEConfig:
    in_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(8, 16, 32, 32), norm_num_groups=4)


# ----------------------------------------------------------------------------- scheduler
class DDIMScheduler:
    """beta_schedule="scaled_linear", 1000 train steps, steps_offset=0 (optimize_token.py:25-34)."""

    def __init__(self, beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps: int):
        ratio = self.num_train_timesteps // num_inference_steps
        self.timesteps = (torch.arange(num_inference_steps, dtype=torch.int64) * ratio).flip(0)

    def add_noise(self, original, noise, timesteps):
        t = torch.as_tensor(timesteps).reshape(-1).to(torch.int64).cpu()
        a = self.alphas_cumprod[t].to(original.device, original.dtype)
        sa = a.sqrt().reshape(-1, *([1] * (original.dim() - 1)))
        sb = (1.0 - a).sqrt().reshape(-1, *([1] * (original.dim() - 1)))
        return sa * original + sb * noise


# ----------------------------------------------------------------------------- attention
class CrossAttention(nn.Module):  # the class NAME is part of the reference's discovery protocol
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        ctx_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(ctx_dim, inner, bias=False)
        self.to_v = nn.Linear(ctx_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def reshape_heads_to_batch_dim(self, t):
        b, s, c = t.shape
        h = self.heads
        return t.reshape(b, s, h, c // h).permute(0, 2, 1, 3).reshape(b * h, s, c // h)

    def reshape_batch_dim_to_heads(self, t):
        bh, s, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def forward(self, hidden_states, context=None, mask=None):
        ctx = hidden_states if context is None else context
        q = self.reshape_heads_to_batch_dim(self.to_q(hidden_states))
        k = self.reshape_heads_to_batch_dim(self.to_k(ctx))
        v = self.reshape_heads_to_batch_dim(self.to_v(ctx))
        scores = torch.baddbmm(
            torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
            q, k.transpose(-1, -2), beta=0, alpha=self.scale)
        probs = scores.softmax(dim=-1)
        out = self.reshape_batch_dim_to_heads(torch.bmm(probs, v))
        return self.to_out[1](self.to_out[0](out))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, cross_attention_dim, heads, dim_head)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, groups):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        res = x
        y = self.proj_in(self.norm(x))
        y = y.permute(0, 2, 3, 1).reshape(b, h * w, -1)
        for blk in self.transformer_blocks:
            y = blk(y, context=context)
        y = y.reshape(b, h, w, -1).permute(0, 3, 1, 2)
        return self.proj_out(y) + res


# ----------------------------------------------------------------------------- resnet / sampling
class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_channels, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, channels, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, cfg: UNetConfig, has_attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        for i in range(cfg.layers_per_block):
            self.resnets.append(ResnetBlock2D(cin if i == 0 else cout, cout, temb, cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(cfg.attention_head_dim, cout // cfg.attention_head_dim,
                                                          cout, cfg.cross_attention_dim, cfg.norm_num_groups))
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, context):
        outs = ()
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context=context)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class _MidBlock(nn.Module):
    def __init__(self, c, temb, cfg: UNetConfig):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(cfg.attention_head_dim, c // cfg.attention_head_dim, c,
                                                            cfg.cross_attention_dim, cfg.norm_num_groups)])

    def forward(self, x, temb, context):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context=context)
        return self.resnets[1](x, temb)


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, temb, cfg: UNetConfig, has_attn, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = cprev if i == 0 else cout
            self.resnets.append(ResnetBlock2D(rin + skip, cout, temb, cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(cfg.attention_head_dim, cout // cfg.attention_head_dim,
                                                          cout, cfg.cross_attention_dim, cfg.norm_num_groups))
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, context):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context=context)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


def timestep_features(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos(t f), sin(t f)]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t.reshape(-1, 1).float() * freqs[None]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, c in enumerate(ch):
            cin, out = out, c
            self.down_blocks.append(_DownBlock(cin, out, temb, cfg, cfg.down_has_attn[i], i != len(ch) - 1))
        self.mid_block = _MidBlock(ch[-1], temb, cfg)
        self.up_blocks = nn.ModuleList()
        rev = tuple(reversed(ch))
        up_has_attn = tuple(reversed(cfg.down_has_attn))
        out = rev[0]
        for i, c in enumerate(rev):
            prev, out = out, c
            cin = rev[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(_UpBlock(cin, out, prev, temb, cfg, up_has_attn[i], i != len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states):
        t = torch.as_tensor(timestep, device=sample.device).reshape(-1)
        if t.numel() == 1:
            t = t.expand(sample.shape[0])
        emb = self.time_embedding(timestep_features(t, self.cfg.block_out_channels[0]))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(self.conv_act(self.conv_norm_out(x)))
        return {"sample": x}


# ----------------------------------------------------------------------------- VAE encoder
class AttentionBlock(nn.Module):
    def __init__(self, channels, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)
        q, k, v = self.query(y), self.key(y), self.value(y)
        s = 1.0 / math.sqrt(math.sqrt(c))
        p = torch.softmax(torch.matmul(q * s, (k * s).transpose(-1, -2)).float(), dim=-1).to(v.dtype)
        y = self.proj_attn(torch.matmul(p, v)).transpose(1, 2).reshape(b, c, h, w)
        return y + x


class _EncDown(nn.Module):
    def __init__(self, cin, cout, n, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, 1e-6) for i in range(n)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class _EncMid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, groups, 1e-6) for _ in range(2)])
        self.attentions = nn.ModuleList([AttentionBlock(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, c in enumerate(ch):
            cin, out = out, c
            self.down_blocks.append(_EncDown(cin, out, cfg.layers_per_block, cfg.norm_num_groups, i != len(ch) - 1))
        self.mid_block = _EncMid(ch[-1], cfg.norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class _LatentDist:
    def __init__(self, moments):
        self.mean, self.logvar = moments.chunk(2, dim=1)


class AutoencoderKL(nn.Module):
    """Encoder half only (the path never decodes: ptp_utils.py:289-304)."""

    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)

    def encode(self, x):
        return {"latent_dist": _LatentDist(self.quant_conv(self.encoder(x)))}


# ----------------------------------------------------------------------------- pipeline facade
class Pipeline:
    """Duck-type of the StableDiffusionPipeline attributes the hot path touches."""

    def __init__(self, unet, vae, scheduler):
        self.unet, self.vae, self.scheduler = unet, vae, scheduler
        self.text_encoder = nn.Identity()  # loaded but never executed on the live path


def make_pipeline(unet_cfg: UNetConfig = UNetConfig(), vae_cfg: VAEConfig = VAEConfig(), seed: int = 0,
                  attn_gain: float = 1.0) -> Pipeline:
    """Seeded synthetic weights (PyTorch default inits), frozen, fp32, CPU.

    ``attn_gain`` multiplies every cross-attention ``to_q`` weight so synthetic weights give
    peaky (trained-looking) attention instead of near-uniform maps.
    """
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    unet = UNet2DConditionModel(unet_cfg)
    vae = AutoencoderKL(vae_cfg)
    torch.random.set_rng_state(g)
    if attn_gain != 1.0:
        with torch.no_grad():
            for name, p in unet.named_parameters():
                if name.endswith("attn2.to_q.weight"):
                    p.mul_(attn_gain)
    for p in list(unet.parameters()) + list(vae.parameters()):
        p.requires_grad_(False)
    unet.eval(); vae.eval()
    sched = DDIMScheduler()
    sched.set_timesteps(50)
    return Pipeline(unet, vae, sched)
