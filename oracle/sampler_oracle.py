"""Torch restatement of the reference EDM Euler sampler, condition assembly and post-process.

TEST INFRASTRUCTURE (see oracle/__init__.py) — never imported by the product.

Follows model_diffusion_renderer.py:16-82 (scheduler), :88-96 (conditioner), :158-235 (conditions, loop)
and diffusion_renderer_pipeline.py:300-318 (post-process).  Pinned against the real reference in
tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .dit_oracle import dit_forward
from .weights import DitDims

SIGMA_DATA = 0.5


def sigma_schedule(num_steps: int, sigma_max: float = 80.0, sigma_min: float = 0.02, device=None) -> torch.Tensor:
    """model_diffusion_renderer.py:23-28 — logspace(80 -> 0.02, N) ++ [0], fp32."""
    s = torch.logspace(np.log10(sigma_max), np.log10(sigma_min), num_steps, device=device, dtype=torch.float32)
    return torch.cat([s, torch.zeros(1, device=device, dtype=torch.float32)])


def scale_model_input(x: torch.Tensor, sigma: torch.Tensor, sigma_data: float = SIGMA_DATA) -> torch.Tensor:
    """model_diffusion_renderer.py:30-44 — fp32 x / sqrt(sigma^2 + sigma_d^2), cast back."""
    c_in = 1 / torch.sqrt(sigma.float() ** 2 + sigma_data ** 2)
    return (x.float() * c_in).to(x.dtype)


def euler_step(F_out: torch.Tensor, sigma: torch.Tensor, sigma_next: torch.Tensor, x: torch.Tensor,
               sigma_data: float = SIGMA_DATA) -> torch.Tensor:
    """model_diffusion_renderer.py:46-82 — all fp32, result cast to x.dtype."""
    s = sigma.float()
    c_skip = sigma_data ** 2 / (s ** 2 + sigma_data ** 2)
    c_out = (s * sigma_data) / torch.sqrt(s ** 2 + sigma_data ** 2)
    xf = x.float()
    den = c_skip * xf + c_out * F_out.float()
    return (xf + (xf - den) / s * (sigma_next - s)).to(x.dtype)


def latent_conditions(data_batch: Dict[str, torch.Tensor], condition_keys: Sequence[str], append_mask: bool,
                      encode: Callable[[torch.Tensor], torch.Tensor], latent_shape) -> torch.Tensor:
    """model_diffusion_renderer.py:158-197 — encode(x)*sigma_d per key (zeros when absent), optional mask channel."""
    B, C, T, H, W = latent_shape
    ref = next(v for v in data_batch.values() if isinstance(v, torch.Tensor) and v.ndim == 5)
    out: List[torch.Tensor] = []
    for key in condition_keys:
        actual = key if key in data_batch else ("rgb" if ("rgb" in data_batch and key == "image") else None)
        if actual is None:
            out.append(torch.zeros(latent_shape, dtype=ref.dtype, device=ref.device))
            if append_mask:
                out.append(torch.zeros((B, 1, T, H, W), dtype=ref.dtype, device=ref.device))
        else:
            z = (encode(data_batch[actual]) * SIGMA_DATA).contiguous()
            out.append(z)
            if append_mask:
                out.append(torch.ones((B, 1, T, H, W), dtype=z.dtype, device=z.device))
    return torch.cat(out, dim=1)


def sample(sd, d: DitDims, latent_condition: torch.Tensor, context_index: Optional[torch.Tensor],
           state_shape, num_steps: int, seed: int, guidance: float = 0.0,
           per_step: Optional[List[torch.Tensor]] = None, noise: Optional[torch.Tensor] = None,
           teacher: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """model_diffusion_renderer.py:211-235 — seed, noise*sigma0, Euler loop, optional CFG.

    `per_step` collects x_t after every step.  `teacher` (a list of x_t, one per step) makes the loop
    teacher-forced: step i starts from teacher[i] instead of its own previous output.
    """
    dtype, device = latent_condition.dtype, latent_condition.device
    sig = sigma_schedule(num_steps, device=device)
    if noise is None:
        torch.manual_seed(seed)
        noise = torch.randn(size=(1, *state_shape), dtype=dtype, device=device)
    x = noise * sig[0]
    for i in range(num_steps):
        if teacher is not None:
            x = teacher[i]
        xin = scale_model_input(x, sig[i])
        Fc = dit_forward(sd, d, xin, sig[i], latent_condition, context_index)
        if guidance > 0:
            Fu = dit_forward(sd, d, xin, sig[i], torch.zeros_like(latent_condition),
                             None if context_index is None else torch.zeros_like(context_index))
            Fc = Fc + guidance * (Fc - Fu)
        x = euler_step(Fc, sig[i], sig[i + 1], x)
        if per_step is not None:
            per_step.append(x.clone())
    return x


def postprocess(video: torch.Tensor, normalize_normal: bool = False) -> np.ndarray:
    """diffusion_renderer_pipeline.py:300-318 — normal blend, [-1,1] -> uint8 BTHWC (truncating cast)."""
    if normalize_normal:
        norm = torch.norm(video, dim=1, p=2, keepdim=True)
        vn = video / norm.clamp(min=1e-12)
        blend = torch.clip((norm - 0.2) / (0.4 - 0.2), 0, 1)
        video = vn * blend + video * (1 - blend)
    video = (1.0 + video).clamp(0, 2) / 2
    video = video.permute(0, 2, 3, 4, 1)
    return (video * 255).to(torch.uint8).cpu().numpy()
