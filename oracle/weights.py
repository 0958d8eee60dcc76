"""Deterministic synthetic weights with the reference's state_dict key names.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference ships no weights
and every BASELINE config is "random-init"; both the oracle and the CUDA path
load the SAME dict produced here, so parity never depends on the init scheme.

Key names / shapes follow the reference modules:
  CleanGeneralDIT.py:241-257 (attention projections + per-head RMSNorm weights),
  :341-347 (t_embedder), :385-387 (x_embedder.proj.1), :445-447 (MLP),
  :484-488 and :558-562 (adaLN-LoRA pairs), :555 (final linear), :652
  (affline_norm), :729 (context_embedding), :91 (pos_embedder.seq buffer);
  model_diffusion_renderer.py:9-14,114-117 (logvar placeholder).
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, asdict
from typing import Dict, Iterator, Tuple

import torch


@dataclass(frozen=True)
class DitDims:
    """Shape parameters of one GeneralDIT instance (diffusion_renderer_config.py:47-103)."""
    model_channels: int = 4096
    num_blocks: int = 28
    num_heads: int = 32
    in_channels: int = 16
    out_channels: int = 16
    additional_concat_ch: int = 16      # 16 inverse (config:168), 136 forward (config:232)
    crossattn_emb_channels: int = 1024
    adaln_lora_dim: int = 256
    mlp_ratio: float = 4.0
    patch_spatial: int = 2
    patch_temporal: int = 1
    use_context_embedding: bool = True  # inverse True (config:169), forward False (config:233)
    concat_padding_mask: bool = True

    @property
    def head_dim(self) -> int:
        return self.model_channels // self.num_heads

    @property
    def hidden(self) -> int:
        return int(self.model_channels * self.mlp_ratio)

    @property
    def in_total(self) -> int:
        return self.in_channels + self.additional_concat_ch + (1 if self.concat_padding_mask else 0)

    @property
    def patch_dim(self) -> int:
        return self.in_total * self.patch_spatial ** 2 * self.patch_temporal

    @property
    def out_patch_dim(self) -> int:
        return self.out_channels * self.patch_spatial ** 2 * self.patch_temporal

    def net_kwargs(self) -> dict:
        """kwargs accepted by the reference CleanDiffusionRendererGeneralDIT (CleanGeneralDIT.py:722)."""
        return dict(
            model_channels=self.model_channels, num_blocks=self.num_blocks, num_heads=self.num_heads,
            in_channels=self.in_channels, out_channels=self.out_channels,
            crossattn_emb_channels=self.crossattn_emb_channels, block_config="FA-CA-MLP",
            mlp_ratio=self.mlp_ratio, patch_spatial=self.patch_spatial, patch_temporal=self.patch_temporal,
            concat_padding_mask=self.concat_padding_mask, affline_emb_norm=True,
            additional_concat_ch=self.additional_concat_ch, use_context_embedding=self.use_context_embedding,
        )

    def asdict(self) -> dict:
        return asdict(self)


TINY_INVERSE = DitDims(model_channels=512, num_blocks=4, num_heads=4)               # BASELINE config 1
TINY_FORWARD = DitDims(model_channels=512, num_blocks=4, num_heads=4, additional_concat_ch=136,
                       use_context_embedding=False)
MICRO_INVERSE = DitDims(model_channels=256, num_blocks=2, num_heads=2)               # golden-vector size
MICRO_FORWARD = DitDims(model_channels=256, num_blocks=2, num_heads=2, additional_concat_ch=136,
                        use_context_embedding=False)
FULL_INVERSE = DitDims()                                                            # BASELINE config 2
FULL_FORWARD = DitDims(additional_concat_ch=136, use_context_embedding=False)       # BASELINE config 3


def net_param_shapes(d: DitDims) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (key, shape, kind) for every entry of the reference *model* state_dict, in module order."""
    D, R, C = d.model_channels, d.adaln_lora_dim, d.crossattn_emb_channels
    yield "net.x_embedder.proj.1.weight", (D, d.patch_dim), "linear"
    yield "net.t_embedder.1.linear_1.weight", (D, D), "linear"
    yield "net.t_embedder.1.linear_2.weight", (3 * D, D), "linear"
    yield "net.pos_embedder.seq", (max(512, d.head_dim),), "arange"
    for i in range(d.num_blocks):
        p = f"net.blocks.block{i}.blocks"
        for j, ctx in ((0, D), (1, C)):
            a = f"{p}.{j}.block.attn"
            yield f"{a}.to_q.0.weight", (D, D), "linear"
            yield f"{a}.to_q.1.weight", (d.head_dim,), "norm"
            yield f"{a}.to_k.0.weight", (D, ctx), "linear"
            yield f"{a}.to_k.1.weight", (d.head_dim,), "norm"
            yield f"{a}.to_v.0.weight", (D, ctx), "linear"
            yield f"{a}.to_out.0.weight", (D, D), "linear"
            yield f"{p}.{j}.adaLN_modulation.1.weight", (R, D), "linear"
            yield f"{p}.{j}.adaLN_modulation.2.weight", (3 * D, R), "linear"
        yield f"{p}.2.block.layer1.weight", (d.hidden, D), "linear"
        yield f"{p}.2.block.layer2.weight", (D, d.hidden), "linear"
        yield f"{p}.2.adaLN_modulation.1.weight", (R, D), "linear"
        yield f"{p}.2.adaLN_modulation.2.weight", (3 * D, R), "linear"
    yield "net.final_layer.linear.weight", (d.out_patch_dim, D), "linear"
    yield "net.final_layer.adaLN_modulation.1.weight", (R, D), "linear"
    yield "net.final_layer.adaLN_modulation.2.weight", (2 * D, R), "linear"
    yield "net.affline_norm.weight", (D,), "norm"
    if d.use_context_embedding:
        yield "net.context_embedding.weight", (16, C), "embed"
    yield "logvar.0.freqs", (128,), "embed"
    yield "logvar.0.phases", (128,), "embed"
    yield "logvar.1.weight", (1, 128), "linear"


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def make_state_dict(d: DitDims, seed: int = 0, dtype: torch.dtype = torch.float32,
                    device: str | torch.device = "cpu", gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Per-tensor seeded init (order-independent, so a subset can be regenerated).

    linear: U(-g/sqrt(fan_in), g/sqrt(fan_in)) (nn.Linear's default bound); norm weights: 1 + 0.1*N(0,1)
    (so the per-head RMSNorm weights are exercised); embeddings: N(0,1).
    Values are drawn in fp32 with a generator on `device`, then cast to `dtype`.
    """
    device = torch.device(device)
    out: Dict[str, torch.Tensor] = {}
    for key, shape, kind in net_param_shapes(d):
        g = torch.Generator(device=device)
        g.manual_seed(_key_seed(seed, key))
        if kind == "arange":
            t = torch.arange(shape[0], dtype=torch.float32, device=device)
            out[key] = t  # buffer stays fp32-valued; cast below like module.to(dtype) would
        elif kind == "linear":
            bound = gain / math.sqrt(shape[-1])
            t = (torch.rand(shape, generator=g, device=device, dtype=torch.float32) * 2 - 1) * bound
        elif kind == "norm":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        else:
            t = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        out[key] = t.to(dtype)
    return out


def net_only(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Strip the 'net.' prefix (keys of CleanDiffusionRendererGeneralDIT.state_dict())."""
    return {k[4:]: v for k, v in sd.items() if k.startswith("net.")}
