"""A tiny deterministic stand-in with the reference tokenizer's call surface (CleanVAE.py:9-67).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The sampler parity tests need *a* tokenizer so that the
reference model's `encode`/`decode` calls (model_diffusion_renderer.py:138-156) run; its arithmetic is
irrelevant to DiT/sampler parity, so this one is a fixed linear pool/unpool: 8x8 spatial mean, causal
1+8k temporal grouping, and a fixed 3<->16 channel mix.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class StubVAE:
    latent_ch = 16
    spatial_compression_factor = 8
    temporal_compression_factor = 8

    def __init__(self):
        i = torch.arange(16, dtype=torch.float32)[:, None]
        j = torch.arange(3, dtype=torch.float32)[None, :]
        self.mix = torch.cos(0.7 * i + 1.3 * j + 0.1) / math.sqrt(3.0)      # (16,3)
        self.config = {"latent_channels": 16, "spatial_compression_ratio": 8, "temporal_compression_ratio": 8}

    def get_latent_num_frames(self, n: int) -> int:
        return 1 if n == 1 else (n - 1) // 8 + 1

    def get_pixel_num_frames(self, n: int) -> int:
        return 1 if n == 1 else (n - 1) * 8 + 1

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim != 5:
            raise ValueError(f"expects a 5D input (B, C, T, H, W), but got {x.shape}")
        B, C, T, H, W = x.shape
        xs = F.avg_pool3d(x.float(), (1, 8, 8))
        first = xs[:, :, :1]
        rest = xs[:, :, 1:]
        if rest.shape[2] > 0:
            rest = rest.reshape(B, C, (T - 1) // 8, 8, H // 8, W // 8).mean(dim=3)
            xs = torch.cat([first, rest], dim=2)
        else:
            xs = first
        z = torch.einsum("oc,bcthw->bothw", self.mix.to(x.device), xs)
        return z.to(x.dtype)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        if z.ndim != 5:
            raise ValueError(f"expects a 5D latent (B, C, T, H, W), but got {z.shape}")
        v = torch.einsum("oc,bothw->bcthw", self.mix.to(z.device), z.float())
        t = self.get_pixel_num_frames(z.shape[2])
        idx = torch.tensor([0] + [1 + (k // 8) for k in range(t - 1)], device=z.device)
        v = v.index_select(2, idx)
        v = v.repeat_interleave(8, dim=3).repeat_interleave(8, dim=4)
        return v.to(z.dtype)

    def to(self, device):
        return self

    def reset_dtype(self, dtype):
        return None
