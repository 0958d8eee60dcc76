"""Functional torch restatement of the reference GeneralDIT forward.

TEST INFRASTRUCTURE (see oracle/__init__.py) — never imported by the product.

It takes the reference's own state_dict (same key names) and recomputes
`CleanDiffusionRendererGeneralDIT.forward` (CleanGeneralDIT.py:731-751) op by
op, in the dtype of the weights, so that a bf16 run rounds at the same points
as the reference does.  It includes the single oracle
patch (heads flattened before `to_out`), without
which the reference cannot run.  Pinned against the real reference modules in
tests/test_oracle_vs_reference.py and tests/golden/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .weights import DitDims

SD = Dict[str, torch.Tensor]


# ---------------------------------------------------------------- small pieces
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """CleanGeneralDIT.py:23-33 — fp32 inside, (x_normed * weight) then cast back."""
    xf = x.float()
    inv = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (xf * inv * weight).to(x.dtype)


def sigma_embedding(sigma: torch.Tensor, channels: int) -> torch.Tensor:
    """CleanGeneralDIT.py:321-335 — [cos(s*w_i), sin(s*w_i)], w_i = exp(-ln(1e4) i/half); returns sigma.dtype."""
    half = channels // 2
    w = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=sigma.device) / (half - 0.0))
    ang = sigma[:, None].float() * w[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(sigma.dtype)


def rope_angles(d: DitDims, T: int, H: int, W: int, dtype: torch.dtype, device,
                seq: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CleanGeneralDIT.py:94-159 — (T*H*W, head_dim) angles in the [t,h,w]*2 layout.

    Quirk kept on purpose: `seq`, `dim_spatial_range` and `dim_temporal_range` are module
    *buffers* (:91,:106-111), so `model.to(bfloat16)` casts them and the whole angle computation (pow,
    reciprocal, outer) then runs in bf16, not only the final `.to(x_patches.dtype)` (:159).
    """
    hd = d.head_dim
    dim_h = hd // 6 * 2
    dim_t = hd - 2 * dim_h
    r_s = (torch.arange(0, dim_h, 2, device=device)[: dim_h // 2].float() / dim_h).to(dtype)
    r_t = (torch.arange(0, dim_t, 2, device=device)[: dim_t // 2].float() / dim_t).to(dtype)
    f_s = 1.0 / (10000.0 ** r_s)              # h and w share theta = 1e4 (ntk factor 1.0, :126-127)
    f_t = 1.0 / ((10000.0 * 2.0) ** r_t)      # temporal ntk factor 2.0 (:116,:128)
    if seq is None:
        seq = torch.arange(max(512, hd), dtype=torch.float32, device=device)
    seq = seq.to(dtype)
    a_t = torch.outer(seq[:T], f_t)[:, None, None, :].expand(T, H, W, -1)
    a_h = torch.outer(seq[:H], f_s)[None, :, None, :].expand(T, H, W, -1)
    a_w = torch.outer(seq[:W], f_s)[None, None, :, :].expand(T, H, W, -1)
    ang = torch.cat([a_t, a_h, a_w, a_t, a_h, a_w], dim=-1)
    return ang.reshape(T * H * W, hd).to(dtype)


def apply_rope(x_sbhd: torch.Tensor, ang: torch.Tensor) -> torch.Tensor:
    """CleanGeneralDIT.py:45-84 — x*cos + rotate_half(x)*sin with cos/sin taken in x.dtype."""
    a = ang[:, None, None, :].expand_as(x_sbhd)
    c, s = a.cos().to(x_sbhd.dtype), a.sin().to(x_sbhd.dtype)
    x1, x2 = x_sbhd.chunk(2, dim=-1)
    return x_sbhd * c + torch.cat((-x2, x1), dim=-1) * s


def patchify(x: torch.Tensor, p: int) -> torch.Tensor:
    """CleanGeneralDIT.py:409-414 — b c t (h m) (w n) -> b t h w (c m n)  (temporal patch 1)."""
    B, C, T, Hh, Ww = x.shape
    x = x.reshape(B, C, T, Hh // p, p, Ww // p, p)
    return x.permute(0, 2, 3, 5, 1, 4, 6).reshape(B, T, Hh // p, Ww // p, C * p * p)


def unpatchify(y: torch.Tensor, B: int, T: int, Hp: int, Wp: int, p: int, C: int) -> torch.Tensor:
    """CleanGeneralDIT.py:709-716 — (B T) (H W) (ph pw pt C) -> B C (T pt) (H ph) (W pw); C fastest."""
    y = y.reshape(B, T, Hp, Wp, p, p, C)
    return y.permute(0, 6, 1, 2, 4, 3, 5).reshape(B, C, T, Hp * p, Wp * p)


# ---------------------------------------------------------------- sub-blocks
def _modulation(sd: SD, prefix: str, emb: torch.Tensor, lora: torch.Tensor) -> torch.Tensor:
    """CleanGeneralDIT.py:483-488,500-501 — W_b (W_a SiLU(emb)) + lora."""
    h = F.linear(F.silu(emb), sd[f"{prefix}.adaLN_modulation.1.weight"])
    return F.linear(h, sd[f"{prefix}.adaLN_modulation.2.weight"]) + lora


def _attention(sd: SD, prefix: str, d: DitDims, x: torch.Tensor, ctx: Optional[torch.Tensor],
               ang: Optional[torch.Tensor]) -> torch.Tensor:
    """CleanGeneralDIT.py:268-306 (+ :181-203 with the head-flatten patch).  x: (S,B,D); ctx: (M,B,Dc) or None."""
    a = f"{prefix}.block.attn"
    src = x if ctx is None else ctx
    q = F.linear(x, sd[f"{a}.to_q.0.weight"])
    k = F.linear(src, sd[f"{a}.to_k.0.weight"])
    v = F.linear(src, sd[f"{a}.to_v.0.weight"])
    H, hd = d.num_heads, d.head_dim
    q = q.reshape(*q.shape[:2], H, hd)
    k = k.reshape(*k.shape[:2], H, hd)
    v = v.reshape(*v.shape[:2], H, hd)
    q = rms_norm(q, sd[f"{a}.to_q.1.weight"])
    k = rms_norm(k, sd[f"{a}.to_k.1.weight"])
    if ctx is None and ang is not None:
        q, k = apply_rope(q, ang), apply_rope(k, ang)
    o = F.scaled_dot_product_attention(q.permute(1, 2, 0, 3), k.permute(1, 2, 0, 3), v.permute(1, 2, 0, 3))
    o = o.permute(2, 0, 1, 3).flatten(2)           # the one oracle patch: (S,B,H,hd) -> (S,B,H*hd)
    return F.linear(o, sd[f"{a}.to_out.0.weight"])


def _mlp(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """CleanGeneralDIT.py:449-462 — Linear, exact-erf GELU, Linear."""
    return F.linear(F.gelu(F.linear(x, sd[f"{prefix}.block.layer1.weight"])), sd[f"{prefix}.block.layer2.weight"])


def sub_block(sd: SD, prefix: str, kind: str, d: DitDims, x: torch.Tensor, emb: torch.Tensor,
              lora: torch.Tensor, ctx: torch.Tensor, ang: torch.Tensor) -> torch.Tensor:
    """CleanGeneralDIT.py:492-517 — AdaLN(shift, scale, gate), block, gated residual."""
    D = d.model_channels
    shift, scale, gate = _modulation(sd, prefix, emb, lora).chunk(3, dim=1)
    xm = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale.unsqueeze(0)) + shift.unsqueeze(0)
    if kind == "fa":
        out = _attention(sd, prefix, d, xm, None, ang)
    elif kind == "ca":
        out = _attention(sd, prefix, d, xm, ctx, None)
    else:
        out = _mlp(sd, prefix, xm)
    return x + gate.unsqueeze(0) * out


# ---------------------------------------------------------------- whole net
def dit_forward(sd: SD, d: DitDims, x: torch.Tensor, sigma: torch.Tensor, latent_condition: torch.Tensor,
                context_index: Optional[torch.Tensor] = None,
                taps: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """CleanGeneralDIT.py:731-751 + :656-718.  `sd` uses net-level keys (no 'net.' prefix).

    x (B,16,T,H,W) already c_in-scaled; sigma 0-d or (B,); returns F (B,16,T,H,W) in x.dtype.
    `taps`, if given, collects the (S,B,D) residual stream after every transformer block.
    """
    sd = {"net." + k: v for k, v in sd.items()} if "x_embedder.proj.1.weight" in sd else sd
    B, _, T, Hh, Ww = x.shape
    D, p = d.model_channels, d.patch_spatial
    # :734-742 — context token: embedding row, or zeros for the forward renderer
    if d.use_context_embedding:
        ctx = F.embedding(context_index.long(), sd["net.context_embedding.weight"])
        if ctx.ndim == 2:
            ctx = ctx.unsqueeze(1)
    else:
        ctx = torch.zeros(B, 1, d.crossattn_emb_channels, device=x.device, dtype=x.dtype)
    # :664-666 — sigma rounded to x.dtype *before* the sinusoid
    t = sigma.to(x.dtype).flatten()
    e = sigma_embedding(t, D)
    lora = F.linear(F.silu(F.linear(e, sd["net.t_embedder.1.linear_1.weight"])), sd["net.t_embedder.1.linear_2.weight"])
    emb = rms_norm(e, sd["net.affline_norm.weight"])
    # :669-678 — cat [x, cond, ones], patchify, Linear (no bias)
    parts = [x, latent_condition]
    if d.concat_padding_mask:
        parts.append(torch.ones(B, 1, T, Hh, Ww, device=x.device, dtype=x.dtype))
    tok = F.linear(patchify(torch.cat(parts, dim=1), p), sd["net.x_embedder.proj.1.weight"])   # (B,T,Hp,Wp,D)
    Hp, Wp = Hh // p, Ww // p
    ang = rope_angles(d, T, Hp, Wp, tok.dtype, x.device, sd.get("net.pos_embedder.seq"))
    h = tok.reshape(B, T * Hp * Wp, D).permute(1, 0, 2)           # (S,B,D)
    ctx = ctx.permute(1, 0, 2)                                    # (M,B,Dc)
    for i in range(d.num_blocks):
        for j, kind in enumerate(("fa", "ca", "mlp")):
            h = sub_block(sd, f"net.blocks.block{i}.blocks.{j}", kind, d, h, emb, lora, ctx, ang)
        if taps is not None:
            taps.append(h)
    # :567-590 — final AdaLN (first 2D columns of lora) + Linear
    hb = h.permute(1, 0, 2).reshape(B * T, Hp * Wp, D)
    m = F.linear(F.linear(F.silu(emb), sd["net.final_layer.adaLN_modulation.1.weight"]),
                 sd["net.final_layer.adaLN_modulation.2.weight"]) + lora[:, : 2 * D]
    shift, scale = m.chunk(2, dim=1)
    shift = shift.repeat_interleave(T, dim=0)
    scale = scale.repeat_interleave(T, dim=0)
    y = F.layer_norm(hb, (D,), eps=1e-6) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
    y = F.linear(y, sd["net.final_layer.linear.weight"])
    return unpatchify(y, B, T, Hp, Wp, p, d.out_channels)
