"""Torch restatement of the Cosmos CV8x8x8 causal video tokenizer (diffusers.AutoencoderKLCosmos).

TEST INFRASTRUCTURE (see oracle/__init__.py) — never imported by the product.

**PARITY UNPINNED.**  The reference only wraps the third-party class (`CleanVAE.py:3,18,50-51,59-60`:
`AutoencoderKLCosmos.from_pretrained(path)`, `encode(x).latent_dist.sample()`, `decode(z).sample`); the arithmetic lives in
`diffusers/models/autoencoders/autoencoder_kl_cosmos.py` (diffusers >= 0.34, `VAE_config.json:3` says 0.34.0.dev0),
which is neither installed nor vendored here, and the reference holds no test or golden vector for it.  This file
restates the published algorithm from `VAE_config.json` (channels `:7-18`, `patch_size 4` `:539`, `haar` `:540`,
`attention_resolutions [32]`, `resolution 1024`, `num_layers 2`, 8x/8x compression) and SURVEY.md Appendix B, with the
upstream module / state_dict key names, and is self-checked only: parameter count (~105.6 M), shape contract (57 -> 8 and
121 -> 16 latent frames, /8 spatial), causality, and Haar -> inverse-Haar identity (tests/test_vae_oracle_cpu.py).
`latents_mean/std` of the config are not applied, as in the reference (`CleanVAE.encode` returns the raw latent).
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class VaeDims:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 16
    encoder_block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    decode_block_out_channels: Tuple[int, ...] = (256, 512, 512, 512)
    attention_resolutions: Tuple[int, ...] = (32,)
    resolution: int = 1024
    num_layers: int = 2
    patch_size: int = 4
    spatial_compression_ratio: int = 8
    temporal_compression_ratio: int = 8

    @property
    def inner_dim(self) -> int:
        return self.in_channels * self.patch_size ** 3


FULL_VAE = VaeDims()
# same topology at 1/8 of the width: every code path (shortcuts, resamplers, both attentions) at CPU-test cost
SMALL_VAE = VaeDims(encoder_block_out_channels=(64, 64, 128, 128), decode_block_out_channels=(64, 128, 128, 128))


# ---------------------------------------------------------------------------------------------- topology
def _resnet_shapes(p: str, cin: int, cout: int) -> Iterator[Tuple[str, Tuple[int, ...]]]:
    yield f"{p}.norm1.norm.weight", (cin,)
    yield f"{p}.norm1.norm.bias", (cin,)
    yield from _convproj_shapes(f"{p}.conv1", cin, cout)
    yield f"{p}.norm2.norm.weight", (cout,)
    yield f"{p}.norm2.norm.bias", (cout,)
    yield from _convproj_shapes(f"{p}.conv2", cout, cout)
    if cin != cout:
        yield f"{p}.conv_shortcut.weight", (cout, cin, 1, 1, 1)
        yield f"{p}.conv_shortcut.bias", (cout,)


def _convproj_shapes(p: str, cin: int, cout: int):
    yield f"{p}.conv_s.weight", (cout, cin, 1, 3, 3)
    yield f"{p}.conv_s.bias", (cout,)
    yield f"{p}.conv_t.weight", (cout, cout, 3, 1, 1)
    yield f"{p}.conv_t.bias", (cout,)


def _attn_shapes(p: str, c: int):
    yield f"{p}.norm.norm.weight", (c,)
    yield f"{p}.norm.norm.bias", (c,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        yield f"{p}.{n}.weight", (c, c, 1, 1, 1)
        yield f"{p}.{n}.bias", (c,)


def _mid_shapes(p: str, c: int):
    yield from _resnet_shapes(f"{p}.resnets.0", c, c)
    yield from _attn_shapes(f"{p}.attentions.0", c)
    yield from _attn_shapes(f"{p}.temp_attentions.0", c)
    yield from _resnet_shapes(f"{p}.resnets.1", c, c)


def _resampler_shapes(p: str, c: int, spatial: bool, temporal: bool, up: bool):
    # Downsample: conv1 spatial (1,3,3)/s2, conv2 temporal (3,1,1)/s2; Upsample: conv1 temporal, conv2 spatial; conv3 1x1x1
    first, second = ((3, 1, 1), (1, 3, 3)) if up else ((1, 3, 3), (3, 1, 1))
    use_first, use_second = (temporal, spatial) if up else (spatial, temporal)
    if use_first:
        yield f"{p}.conv1.weight", (c, c, *first)
        yield f"{p}.conv1.bias", (c,)
    if use_second:
        yield f"{p}.conv2.weight", (c, c, *second)
        yield f"{p}.conv2.bias", (c,)
    yield f"{p}.conv3.weight", (c, c, 1, 1, 1)
    yield f"{p}.conv3.bias", (c,)


def encoder_plan(d: VaeDims) -> List[dict]:
    """down blocks: channels and resampling flags (autoencoder_kl_cosmos.CosmosEncoder3d.__init__)"""
    ch = d.encoder_block_out_channels
    n_sp = int(math.log2(d.spatial_compression_ratio)) - int(math.log2(d.patch_size))
    n_tp = int(math.log2(d.temporal_compression_ratio)) - int(math.log2(d.patch_size))
    plan = []
    for i in range(len(ch) - 1):
        sp = tp = False
        if i < len(ch) - 2:
            sp, tp = i < n_sp, i < n_tp
        plan.append(dict(cin=ch[i], cout=ch[i + 1], spatial=sp, temporal=tp))
    return plan


def decoder_plan(d: VaeDims) -> List[dict]:
    ch = list(reversed(d.decode_block_out_channels))
    n_sp = int(math.log2(d.spatial_compression_ratio)) - int(math.log2(d.patch_size))
    n_tp = int(math.log2(d.temporal_compression_ratio)) - int(math.log2(d.patch_size))
    plan = []
    for i in range(len(ch) - 1):
        sp = tp = False
        if i < len(ch) - 2:
            sp, tp = 0 < i < n_sp + 1, 0 < i < n_tp + 1
        plan.append(dict(cin=ch[i], cout=ch[i + 1], spatial=sp, temporal=tp))
    return plan


def vae_param_shapes(d: VaeDims) -> Iterator[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of every entry of AutoencoderKLCosmos.state_dict(), module order.  Note: the resolutions visited
    are 256,128,64 (encoder) / 64,128,256 (decoder), never 32, so only the mid blocks carry attention."""
    ech = d.encoder_block_out_channels
    yield from _convproj_shapes("encoder.conv_in", d.inner_dim, ech[0])
    for i, b in enumerate(encoder_plan(d)):
        for j in range(d.num_layers):
            yield from _resnet_shapes(f"encoder.down_blocks.{i}.resnets.{j}", b["cin"] if j == 0 else b["cout"], b["cout"])
        if b["spatial"] or b["temporal"]:
            yield from _resampler_shapes(f"encoder.down_blocks.{i}.downsamplers.0", b["cout"], b["spatial"], b["temporal"], up=False)
    yield from _mid_shapes("encoder.mid_block", ech[-1])
    yield "encoder.norm_out.norm.weight", (ech[-1],)
    yield "encoder.norm_out.norm.bias", (ech[-1],)
    yield from _convproj_shapes("encoder.conv_out", ech[-1], d.latent_channels)
    yield "quant_conv.weight", (d.latent_channels, d.latent_channels, 1, 1, 1)
    yield "quant_conv.bias", (d.latent_channels,)
    yield "post_quant_conv.weight", (d.latent_channels, d.latent_channels, 1, 1, 1)
    yield "post_quant_conv.bias", (d.latent_channels,)
    dch = list(reversed(d.decode_block_out_channels))
    yield from _convproj_shapes("decoder.conv_in", d.latent_channels, dch[0])
    yield from _mid_shapes("decoder.mid_block", dch[0])
    for i, b in enumerate(decoder_plan(d)):
        for j in range(d.num_layers + 1):
            yield from _resnet_shapes(f"decoder.up_blocks.{i}.resnets.{j}", b["cin"] if j == 0 else b["cout"], b["cout"])
        if b["spatial"] or b["temporal"]:
            yield from _resampler_shapes(f"decoder.up_blocks.{i}.upsamplers.0", b["cout"], b["spatial"], b["temporal"], up=True)
    yield "decoder.norm_out.norm.weight", (dch[-1],)
    yield "decoder.norm_out.norm.bias", (dch[-1],)
    yield from _convproj_shapes("decoder.conv_out", dch[-1], d.inner_dim)


def make_vae_state_dict(d: VaeDims, seed: int = 0, dtype=torch.float32, device="cpu") -> SD:
    """Deterministic per-key init: conv weights U(+-1/sqrt(fan_in)), biases 0.05 N(0,1), norm weights 1 + 0.1 N(0,1)."""
    out: SD = {}
    for key, shape in vae_param_shapes(d):
        g = torch.Generator(device=device)
        g.manual_seed(int.from_bytes(hashlib.sha256(f"vae{seed}:{key}".encode()).digest()[:7], "little"))
        if len(shape) == 5:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            t = (torch.rand(shape, generator=g, device=device) * 2 - 1) / math.sqrt(fan_in)
        elif key.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            t = 0.05 * torch.randn(shape, generator=g, device=device)
        out[key] = t.to(dtype)
    return out


# ---------------------------------------------------------------------------------------------- primitives
def causal_conv3d(sd: SD, p: str, x: torch.Tensor, stride=(1, 1, 1), padding: int = 0) -> torch.Tensor:
    """CosmosCausalConv3d: replicate the first frame (k_t - 1) + (1 - stride_t) times in front, zero-pad H and W."""
    w, b = sd[f"{p}.weight"], sd[f"{p}.bias"]
    tpad = (w.shape[2] - 1) + (1 - stride[0])
    if tpad > 0:
        x = torch.cat([x[:, :, :1].repeat(1, 1, tpad, 1, 1), x], dim=2)
    if padding:
        x = F.pad(x, (padding, padding, padding, padding, 0, 0))
    return F.conv3d(x, w, b, stride=stride)


def conv_projection(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """CosmosConvProjection3d: (1,3,3) spatial conv then (3,1,1) causal temporal conv."""
    return causal_conv3d(sd, f"{p}.conv_t", causal_conv3d(sd, f"{p}.conv_s", x, padding=1))


def causal_group_norm(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """CosmosCausalGroupNorm(num_groups=1): GroupNorm over (C,H,W) of every frame separately, eps 1e-6, affine."""
    B, C, T, H, W = x.shape
    y = F.group_norm(x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W), 1, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-6)
    return y.reshape(B, T, C, H, W).permute(0, 2, 1, 3, 4)


def resnet_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    res = causal_conv3d(sd, f"{p}.conv_shortcut", x) if f"{p}.conv_shortcut.weight" in sd else x
    h = conv_projection(sd, f"{p}.conv1", F.silu(causal_group_norm(sd, f"{p}.norm1", x)))
    h = conv_projection(sd, f"{p}.conv2", F.silu(causal_group_norm(sd, f"{p}.norm2", h)))
    return h + res


def attention(sd: SD, p: str, x: torch.Tensor, temporal: bool) -> torch.Tensor:
    """CosmosCausalAttention with one head of dim C: spatial = tokens H*W per frame, no mask; temporal = tokens T per
    pixel with a lower-triangular (causal) mask."""
    B, C, T, H, W = x.shape
    h = causal_group_norm(sd, f"{p}.norm", x)
    q, k, v = (causal_conv3d(sd, f"{p}.{n}", h) for n in ("to_q", "to_k", "to_v"))
    if temporal:
        def tok(t):  # [B,C,T,H,W] -> [B*H*W, 1, T, C]
            return t.permute(0, 3, 4, 2, 1).reshape(B * H * W, 1, T, C)
        mask = torch.tril(torch.ones(T, T, device=x.device)).bool()
        o = F.scaled_dot_product_attention(tok(q), tok(k), tok(v), attn_mask=mask)
        o = o.reshape(B, H, W, T, C).permute(0, 4, 3, 1, 2)
    else:
        def tok(t):  # [B,C,T,H,W] -> [B*T, 1, H*W, C]
            return t.permute(0, 2, 3, 4, 1).reshape(B * T, 1, H * W, C)
        o = F.scaled_dot_product_attention(tok(q), tok(k), tok(v))
        o = o.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)
    return causal_conv3d(sd, f"{p}.to_out.0", o.type_as(q)) + x


def mid_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    x = resnet_block(sd, f"{p}.resnets.0", x)
    x = attention(sd, f"{p}.attentions.0", x, temporal=False)
    x = attention(sd, f"{p}.temp_attentions.0", x, temporal=True)
    return resnet_block(sd, f"{p}.resnets.1", x)


def downsample(sd: SD, p: str, x: torch.Tensor, spatial: bool, temporal: bool) -> torch.Tensor:
    if spatial:
        x = F.pad(x, (0, 1, 0, 1, 0, 0))
        x = causal_conv3d(sd, f"{p}.conv1", x, stride=(1, 2, 2)) + F.avg_pool3d(x, (1, 2, 2), (1, 2, 2))
    if temporal:
        x = torch.cat([x[:, :, :1], x], dim=2)
        x = causal_conv3d(sd, f"{p}.conv2", x, stride=(2, 1, 1)) + F.avg_pool3d(x, (2, 1, 1), (2, 1, 1))
    return causal_conv3d(sd, f"{p}.conv3", x)


def upsample(sd: SD, p: str, x: torch.Tensor, spatial: bool, temporal: bool) -> torch.Tensor:
    if temporal:
        f = 2 if x.shape[2] > 1 else 1
        x = x.repeat_interleave(f, dim=2)[:, :, f - 1:]
        x = causal_conv3d(sd, f"{p}.conv1", x) + x
    if spatial:
        x = x.repeat_interleave(2, dim=3).repeat_interleave(2, dim=4)
        x = causal_conv3d(sd, f"{p}.conv2", x, padding=1) + x
    return causal_conv3d(sd, f"{p}.conv3", x)


# ---------------------------------------------------------------------------------------------- Haar patching
_S = 0.7071067811865476


def _dwt_level(x: torch.Tensor) -> torch.Tensor:
    """One 3-D Haar level: low = (a+b)/sqrt2, high = (a-b)/sqrt2 along T, then H, then W; sub-bands concatenated
    along C in the order lll, llh, lhl, lhh, hll, hlh, hhl, hhh (time band first), then / sqrt(8)."""
    def split(t, dim):
        a, b = t.unfold(dim, 2, 2).unbind(-1)
        return (a + b) * _S, (a - b) * _S
    xl, xh = split(x, 2)
    out = []
    for xt in (xl, xh):
        for xs in split(xt, 3):
            out.extend(split(xs, 4))
    return torch.cat(out, dim=1) / math.sqrt(8.0)


def _idwt_level(x: torch.Tensor) -> torch.Tensor:
    def merge(lo, hi, dim):
        a, b = (lo + hi) * _S, (lo - hi) * _S
        return torch.stack([a, b], dim=dim + 1).flatten(dim, dim + 1)
    lll, llh, lhl, lhh, hll, hlh, hhl, hhh = torch.chunk(x, 8, dim=1)
    ll, lh, hl, hh = merge(lll, llh, 4), merge(lhl, lhh, 4), merge(hll, hlh, 4), merge(hhl, hhh, 4)
    lo, hi = merge(ll, lh, 3), merge(hl, hh, 3)
    return merge(lo, hi, 2) * math.sqrt(8.0)


def haar_patch(x: torch.Tensor, patch_size: int = 4) -> torch.Tensor:
    """CosmosPatchEmbed3d: first frame repeated `patch_size` times (T: 1+8k -> 4+8k), then log2(p) DWT levels."""
    x = torch.cat([x[:, :, :1].repeat_interleave(patch_size, dim=2), x[:, :, 1:]], dim=2)
    for _ in range(int(math.log2(patch_size))):
        x = _dwt_level(x)
    return x


def haar_unpatch(x: torch.Tensor, patch_size: int = 4) -> torch.Tensor:
    for _ in range(int(math.log2(patch_size))):
        x = _idwt_level(x)
    return x[:, :, patch_size - 1:]


# ---------------------------------------------------------------------------------------------- encoder / decoder
def encode(sd: SD, d: VaeDims, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKLCosmos.encode(x).latent_dist.sample() — the identity distribution returns the moments unchanged."""
    h = conv_projection(sd, "encoder.conv_in", haar_patch(x, d.patch_size))
    for i, b in enumerate(encoder_plan(d)):
        for j in range(d.num_layers):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h)
        if b["spatial"] or b["temporal"]:
            h = downsample(sd, f"encoder.down_blocks.{i}.downsamplers.0", h, b["spatial"], b["temporal"])
    h = mid_block(sd, "encoder.mid_block", h)
    h = conv_projection(sd, "encoder.conv_out", F.silu(causal_group_norm(sd, "encoder.norm_out", h)))
    return causal_conv3d(sd, "quant_conv", h)


def decode(sd: SD, d: VaeDims, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKLCosmos.decode(z).sample"""
    h = conv_projection(sd, "decoder.conv_in", causal_conv3d(sd, "post_quant_conv", z))
    h = mid_block(sd, "decoder.mid_block", h)
    for i, b in enumerate(decoder_plan(d)):
        for j in range(d.num_layers + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h)
        if b["spatial"] or b["temporal"]:
            h = upsample(sd, f"decoder.up_blocks.{i}.upsamplers.0", h, b["spatial"], b["temporal"])
    h = conv_projection(sd, "decoder.conv_out", F.silu(causal_group_norm(sd, "decoder.norm_out", h)))
    return haar_unpatch(h, d.patch_size)


class OracleVAE:
    """The reference tokenizer surface (CleanVAE.py:9-67) on top of the restatement — a checker for the B200 tokenizer."""
    latent_ch = 16
    spatial_compression_factor = 8
    temporal_compression_factor = 8

    def __init__(self, sd: SD, dims: VaeDims = FULL_VAE):
        self.sd, self.dims = sd, dims
        self.config = {"latent_channels": 16, "spatial_compression_ratio": 8, "temporal_compression_ratio": 8}

    def get_latent_num_frames(self, n: int) -> int:
        return 1 if n == 1 else (n - 1) // 8 + 1

    def get_pixel_num_frames(self, n: int) -> int:
        return 1 if n == 1 else (n - 1) * 8 + 1

    @torch.no_grad()
    def encode(self, x):
        if x.ndim != 5:
            raise ValueError(f"expects a 5D input (B, C, T, H, W), but got {x.shape}")
        return encode(self.sd, self.dims, x)

    @torch.no_grad()
    def decode(self, z):
        if z.ndim != 5:
            raise ValueError(f"expects a 5D latent (B, C, T, H, W), but got {z.shape}")
        return decode(self.sd, self.dims, z)

    def to(self, device):
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    def reset_dtype(self, dtype):
        self.sd = {k: v.to(dtype) for k, v in self.sd.items()}
