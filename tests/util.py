"""Shared helpers of the test suite (test infrastructure: may import oracle/)."""
from __future__ import annotations

import torch

from oracle.weights import DitDims, make_state_dict


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def psnr_u8(a, b) -> float:
    import numpy as np
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean()
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def model_config(dims: DitDims, model_type: str, height: int = 64, width: int = 96, frames: int = 9) -> dict:
    """A product/reference model config for arbitrary (small) GeneralDIT dims."""
    from drb200 import diffusion_renderer_config as cfg
    c = cfg.get_config_by_model_type(model_type, height, width, frames)
    c["model_type"] = model_type
    c["net"].update(dims.net_kwargs())
    c["net"]["adaln_lora_dim"] = dims.adaln_lora_dim
    return c


def build_product_model(dims: DitDims, model_type: str, seed: int, device="cuda", dtype=torch.bfloat16, vae=None):
    """CleanDiffusionRendererModel of the product package carrying the oracle's deterministic weights."""
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    model = CleanDiffusionRendererModel(model_config(dims, model_type))
    sd = make_state_dict(dims, seed=seed, dtype=torch.float32)
    model.load_state_dict(sd, strict=True)
    model = model.to(device=device, dtype=dtype)
    model.vae = vae
    return model, {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}
