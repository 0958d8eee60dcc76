"""EDM sampler + conditioning model of DiffusionRenderer on B200 — drop-in for the reference
`model_diffusion_renderer.py` (same classes, method names, argument meaning and error behaviour).

The Euler loop of `generate_samples_from_batch` (reference :211-235) runs entirely as sm_100a kernels on the current
stream: per step one fused c_in-scale + patchify, the GeneralDIT kernels, and one fused unpatchify + CFG + Euler
update; the sigma schedule sits in device memory so no step synchronises with the host.  The noise draw stays in
PyTorch (`torch.manual_seed(seed)` + `torch.randn`, :216,:222) so that the same seed gives the same noise as the
reference on the same device.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .CleanGeneralDIT import BF16, CleanDiffusionRendererGeneralDIT
from .diffusion_renderer_config import get_inverse_renderer_config


class FourierFeaturesPlaceholder(nn.Module):
    """Only so that the checkpoint keys `logvar.0.{freqs,phases}` load (reference :9-14)."""

    def __init__(self, num_channels, **kwargs):
        super().__init__()
        self.register_buffer("freqs", torch.randn(num_channels))
        self.register_buffer("phases", torch.randn(num_channels))

    def forward(self, x):
        return x


class _StepResult:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


class CleanEDMEulerScheduler:
    """sigma schedule and the two EDM formulas (reference :16-82); tensor math is done by drb_edm_* kernels."""

    def __init__(self, sigma_max=80.0, sigma_min=0.02, sigma_data=0.5, **kwargs):
        if sigma_data != 0.5:
            raise ValueError("the kernels are specialised for sigma_data = 0.5 (every reference config)")
        self.sigma_max, self.sigma_min, self.sigma_data = sigma_max, sigma_min, sigma_data
        self.sigmas, self.timesteps, self.current_step = None, None, 0

    def set_timesteps(self, num_steps, device=None):
        # logspace(80 -> 0.02, N) ++ [0] in fp32 — not Karras-rho (reference :23-28)
        sig = torch.logspace(np.log10(self.sigma_max), np.log10(self.sigma_min), num_steps, device=device, dtype=torch.float32)
        self.sigmas = torch.cat([sig, torch.zeros(1, device=device, dtype=torch.float32)])
        self.timesteps = self.sigmas[:-1]
        self.current_step = 0

    def _sigma_dev(self, timestep, like: Tensor) -> Tensor:
        return torch.as_tensor(timestep, dtype=torch.float32).reshape(-1)[:1].to(like.device).contiguous()

    def scale_model_input(self, sample: Tensor, timestep) -> Tensor:
        return ops.edm_scale_input(sample, self._sigma_dev(timestep, sample))

    def step(self, model_output: Tensor, timestep, sample: Tensor) -> _StepResult:
        if self.sigmas is None or self.current_step >= len(self.sigmas) - 1:
            raise RuntimeError("Scheduler not initialized or timesteps exhausted")
        nxt = self.sigmas[self.current_step + 1: self.current_step + 2].to(sample.device).contiguous()
        out = ops.edm_euler_step(model_output, sample, self._sigma_dev(timestep, sample), nxt)
        self.current_step += 1
        return _StepResult(out)


class CleanCondition:
    def __init__(self, **kwargs):
        self.data = kwargs

    def to_dict(self):
        return self.data


class CleanConditioner:
    """cond = the keys present; uncond = zeros of the same shape (reference :88-96)."""

    def get_condition_uncondition(self, data_batch: Dict) -> Tuple[CleanCondition, CleanCondition]:
        cond, uncond = {}, {}
        for key in ("latent_condition", "context_index"):
            if key in data_batch:
                cond[key] = data_batch[key]
                uncond[key] = torch.zeros_like(data_batch[key])
        return CleanCondition(**cond), CleanCondition(**uncond)


class CleanDiffusionRendererModel(nn.Module):
    def __init__(self, config: Dict[str, Any] = None):
        super().__init__()
        if config is None:
            config = get_inverse_renderer_config()
        self.config = config
        net_config = dict(config.get("net", {}))
        scheduler_config = dict(config.get("scheduler", {}))
        scheduler_config.pop("prediction_type", None)
        self.scheduler = CleanEDMEulerScheduler(**scheduler_config)
        self.conditioner = CleanConditioner()
        self.net = CleanDiffusionRendererGeneralDIT(**net_config)
        self.vae = None
        self.logvar = torch.nn.Sequential(FourierFeaturesPlaceholder(num_channels=128), torch.nn.Linear(128, 1, bias=False))
        model_type = config.get("model_type", "inverse")
        if model_type == "inverse":
            self.condition_keys = config.get("condition_keys", ["image", "rgb"])
        else:
            self.condition_keys = config.get("condition_keys", ["depth", "normal", "basecolor", "roughness", "metallic"])
        self.condition_drop_rate = config.get("condition_drop_rate", 0.0)
        self.append_condition_mask = config.get("append_condition_mask", True)
        self.input_data_key = config.get("input_data_key", "video")
        self.tokenizer = None
        self.eval()

    def _get_tensor_kwargs(self):
        try:
            param = next(self.parameters())
            return {"device": param.device, "dtype": param.dtype}
        except StopIteration:
            return {"device": torch.device("cuda"), "dtype": torch.bfloat16}

    # ---------------------------------------------------------------- tokenizer pass-throughs (reference :138-156)
    def encode(self, x: Tensor) -> Tensor:
        if self.vae is None:
            raise RuntimeError("VAE not initialized in model.")
        if x.ndim != 5:
            raise ValueError(f"Model encode expects a 5D tensor (B,C,T,H,W), but got {x.ndim}D.")
        if hasattr(self.vae, "encode_scaled"):     # B200 tokenizer: the sigma_data factor rides on its output layout kernel
            return self.vae.encode_scaled(x, self.scheduler.sigma_data)
        return self.vae.encode(x) * self.scheduler.sigma_data

    def decode(self, x: Tensor) -> Tensor:
        if self.vae is None:
            raise RuntimeError("VAE not initialized in model.")
        if x.ndim != 5:
            raise ValueError(f"Model decode expects a 5D latent (B,C,T,H,W), but got {x.ndim}D.")
        if hasattr(self.vae, "decode_scaled"):
            return self.vae.decode_scaled(x, 1.0 / self.scheduler.sigma_data)
        return self.vae.decode(x / self.scheduler.sigma_data)

    def decode_u8(self, x: Tensor, normalize_normal: bool = False) -> Optional[Tensor]:
        """decode (:148-156) + the pipeline's post-process (diffusion_renderer_pipeline.py:300-318) as one pass when the
        tokenizer offers it: uint8 (B,T,H,W,3) on the device, or None (the caller then decodes and post-processes)"""
        if self.vae is None:
            raise RuntimeError("VAE not initialized in model.")
        if x.ndim != 5:
            raise ValueError(f"Model decode expects a 5D latent (B,C,T,H,W), but got {x.ndim}D.")
        if not hasattr(self.vae, "decode_u8"):
            return None
        return self.vae.decode_u8(x, 1.0 / self.scheduler.sigma_data, normalize_normal)

    # ---------------------------------------------------------------- conditions (reference :158-209)
    def prepare_diffusion_renderer_latent_conditions(self, data_batch: Dict[str, Tensor], condition_keys: list = None,
                                                     **kwargs) -> Tensor:
        if self.vae is None:
            raise RuntimeError("VAE not initialized in model.")
        if condition_keys is None:
            condition_keys = self.condition_keys
        latent_shape = None
        for key in condition_keys:
            if key in data_batch:
                B, C, T, H, W = data_batch[key].shape
                latent_shape = (B, self.vae.latent_ch, self.vae.get_latent_num_frames(T),
                                H // self.vae.spatial_compression_factor, W // self.vae.spatial_compression_factor)
                break
        if latent_shape is None:
            raise ValueError(f"Could not determine latent shape from keys {condition_keys}.")
        ref = data_batch[self.input_data_key]
        mask_shape = (latent_shape[0], 1, *latent_shape[2:])
        parts = []
        encoded: Dict[int, Tensor] = {}   # the same clip under two keys is encoded once
        for cond_key in condition_keys:
            actual = cond_key if cond_key in data_batch else ("rgb" if "rgb" in data_batch and cond_key == "image" else None)
            if actual is None:
                parts.append(torch.zeros(latent_shape, dtype=ref.dtype, device=ref.device))
                if self.append_condition_mask:
                    parts.append(torch.zeros(mask_shape, dtype=ref.dtype, device=ref.device))
            else:
                src = data_batch[actual]
                state = encoded.get(id(src))
                if state is None:
                    state = encoded[id(src)] = self.encode(src).contiguous()
                parts.append(state)
                if self.append_condition_mask:
                    parts.append(torch.ones(mask_shape, dtype=state.dtype, device=state.device))
        return torch.cat(parts, dim=1)

    def _get_conditions(self, data_batch: Dict, is_negative_prompt: bool = False):
        for key in ("rgb", "basecolor", "normal", "depth", "roughness", "metallic", "image"):
            if key in data_batch:
                self.input_data_key = key
                break
        with torch.no_grad():
            latent_condition = self.prepare_diffusion_renderer_latent_conditions(data_batch, self.condition_keys)
        data_batch["latent_condition"] = latent_condition
        return self.conditioner.get_condition_uncondition(data_batch)

    # ---------------------------------------------------------------- the sampler (reference :211-235)
    def generate_samples_from_batch(self, data_batch: Dict, guidance: float = 0.0, seed: int = 1000,
                                    state_shape: Tuple = None, num_steps: int = 15, **kwargs) -> Tensor:
        with torch.no_grad():
            torch.manual_seed(seed)
            condition, uncondition = self._get_conditions(data_batch)
            tk = self._get_tensor_kwargs()
            self.scheduler.set_timesteps(num_steps, device=tk["device"])
            xt = torch.randn(size=(1, *state_shape), **tk) * self.scheduler.sigmas[0]
            return self.sample_latent(xt, condition.to_dict(), uncondition.to_dict() if guidance > 0 else None,
                                      guidance=guidance, per_step=kwargs.get("per_step"), teacher=kwargs.get("teacher"))

    def generate_samples_multi(self, data_batch: Dict, context_indices, guidance: float = 0.0, seed: int = 1000,
                               state_shape: Tuple = None, num_steps: int = 15, latent_condition: Optional[Tensor] = None) -> Tensor:
        """N passes over ONE clip that differ only in `context_index` (the inverse node's five G-buffer passes,
        nodes.py:187-205), sampled together: same seed, same noise and same encoded condition for every pass — exactly what
        N calls of generate_samples_from_batch compute (:211-235) — but one batched transformer pass per step.
        Returns (N,16,F,h,w).  `latent_condition`: the already encoded clip, when the caller has it."""
        with torch.no_grad():
            torch.manual_seed(seed)
            if latent_condition is None:
                self._get_conditions(data_batch)
                latent_condition = data_batch["latent_condition"]
            tk = self._get_tensor_kwargs()
            conds, unconds = [], []
            for idx in context_indices:
                ci = torch.full((1, 1), int(idx), dtype=torch.long, device=tk["device"])
                c, u = self.conditioner.get_condition_uncondition({"latent_condition": latent_condition, "context_index": ci})
                conds.append(c.to_dict())
                unconds.append(u.to_dict())
            self.scheduler.set_timesteps(num_steps, device=tk["device"])
            xt = torch.randn(size=(1, *state_shape), **tk) * self.scheduler.sigmas[0]
            return self.sample_latent(xt, conds, unconds if guidance > 0 else None, guidance=guidance)

    def sample_latent(self, xt: Tensor, cond, uncond=None, guidance: float = 0.0,
                      per_step: Optional[list] = None, teacher: Optional[list] = None) -> Tensor:
        """Euler loop on device.  `xt` (1,16,F,h,w) bf16 = noise * sigma_0; scheduler.set_timesteps must have run.

        `cond` / `uncond` are the condition dicts of ONE pass (reference :224-234) or lists of N dicts: N passes over the
        same clip (the five G-buffer passes of the inverse node differ only in `context_index`) then run as ONE batched
        transformer pass per step — N sequences stacked along the token-row axis, 2N under CFG (cond and uncond batched,
        SURVEY.md §8f.1) — instead of N (2N) sequential forwards: every row is computed exactly as in the sequential
        form (bit-identical results), the weights stream once per step, and under context parallelism the GEMM grids
        have no short last wave.  Returns (N,16,F,h,w); `xt` may be (1,...) (the same noise for every pass: the node
        seeds every pass identically, nodes.py:187-205) or (N,...).
        `per_step` collects x_t after every step; `teacher` (one x_t per step) makes the loop teacher-forced — both are
        parity-test hooks."""
        net, sig = self.net, self.scheduler.sigmas
        if xt.dtype != BF16 or not xt.is_cuda:
            raise ValueError("the B200 sampler needs a bfloat16 CUDA latent (no CPU fallback)")
        conds = list(cond) if isinstance(cond, (list, tuple)) else [cond]
        N = len(conds)
        cfg = guidance > 0 and uncond is not None
        unconds = []
        if cfg:
            unconds = list(uncond) if isinstance(uncond, (list, tuple)) else [uncond]
            if len(unconds) != N:
                raise ValueError("one unconditional dict per conditional dict")
        if xt.ndim != 5 or xt.shape[0] not in (1, N):
            raise ValueError(f"xt must be (1 or {N}, C, T, H, W), got {tuple(xt.shape)}")
        T, H, W = net._check_input(xt[:1])
        net._ensure_packed()
        # context parallelism (net.enable_context_parallel): this rank owns latent frames [t0, t1); the Euler update is
        # token-local, so the latent stays sharded through the whole loop and is gathered once at the end
        cp = net._cp
        t0, t1 = 0, T
        if cp is not None:
            from .context_parallel import shard_frames
            t0, t1 = shard_frames(T, cp.rank, cp.world)
        Tl = t1 - t0
        seqs = conds + unconds
        ws = net._workspace(Tl, H, W, xt.device, cp, batch=len(seqs))
        sig = sig.to(device=xt.device, dtype=torch.float32).contiguous()

        def local(t5: Optional[Tensor]) -> Optional[Tensor]:
            return None if t5 is None else t5[:, :, t0:t1]

        def full(x_local: Tensor) -> Tensor:
            return x_local if cp is None else cp.all_gather_frames(x_local)

        # constants of the passes, hoisted out of the step loop
        use_ca = False
        for b, c in enumerate(seqs):
            net.prepare_condition(ws, local(c.get("latent_condition")), Tl, H, W, b)
            use_ca = net.prepare_context(ws, net.context_token(c.get("context_index")), b)
        x = local(xt).expand(N, -1, -1, -1, -1).contiguous().clone()
        n = sig.numel() - 1
        for i in range(n):
            if teacher is not None:
                x = local(teacher[i]).expand(N, -1, -1, -1, -1).contiguous().clone()
            net.denoise_step(ws, x, sig[i:i + 1], sig[i + 1:i + 2], use_ca, guidance=guidance if cfg else 0.0)
            if per_step is not None:
                per_step.append(full(x).clone())
        return full(x)
