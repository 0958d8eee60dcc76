"""Tensor-level wrappers over the C ABI: they validate device/dtype/contiguity, pull raw pointers and the current CUDA
stream out of torch, and call libdrb200.so.  PyTorch is used for memory and streams only — no math happens here."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=BF16) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")


def _rows2d(t: torch.Tensor, name: str) -> int:
    """row pitch (elements) of a 2-D tensor whose last dim is contiguous"""
    if t.ndim != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D with a contiguous last dimension")
    return t.stride(0)


def gemm(a: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None, epilogue: int = _lib.EPI_STORE,
         resid: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, cta_group: int = 0, sync=None) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T).  `a`, `w`, `out`, `resid` may be row-strided views.
    `sync`: a drb_cp_sync descriptor (context_parallel.ContextParallel.sync) when `a` is filled by peer GPUs."""
    _req(a, "a"), _req(w, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"shape mismatch: a {tuple(a.shape)} vs w {tuple(w.shape)}")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=BF16)
    _req(out, "out")
    if tuple(out.shape) != (M, N):
        raise ValueError("out has the wrong shape")
    ldr = 0
    if resid is not None:
        _req(resid, "resid"), _req(gate, "gate")
        ldr = _rows2d(resid, "resid")
    if sync is None:
        _lib.call("drb_gemm_bf16", a.data_ptr(), _rows2d(a, "a"), w.data_ptr(), _rows2d(w, "w"), out.data_ptr(),
                  _rows2d(out, "out"), M, N, K, epilogue, _ptr(resid), ldr, _ptr(gate), cta_group, _stream())
    else:
        _lib.call("drb_gemm_bf16_sync", a.data_ptr(), _rows2d(a, "a"), w.data_ptr(), _rows2d(w, "w"), out.data_ptr(),
                  _rows2d(out, "out"), M, N, K, epilogue, _ptr(resid), ldr, _ptr(gate), cta_group, ctypes.byref(sync), _stream())
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int,
              out: Optional[torch.Tensor] = None, max_abs_logit: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [Sq, H*128], k/v [Skv, H*128] (row-strided views of one qkv buffer are fine, same pitch) -> o [Sq, H*128].
    `max_abs_logit`: fp32 device scalar, the caller's bound on |q.k|/sqrt(128); None = unknown (per-tile-max softmax)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, n)
    ld = _rows2d(q, "q")
    if _rows2d(k, "k") != ld or _rows2d(v, "v") != ld:
        raise ValueError("q, k, v must share one row pitch")
    if out is None:
        out = torch.empty((q.shape[0], num_heads * 128), device=q.device, dtype=BF16)
    _req(out, "out")
    _lib.call("drb_attention_bf16_bounded", q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, out.data_ptr(), _rows2d(out, "out"),
              q.shape[0], k.shape[0], num_heads, _bound_ptr(max_abs_logit), _stream())
    return out


def _bound_ptr(b: Optional[torch.Tensor]) -> Optional[int]:
    if b is None:
        return None
    _req(b, "max_abs_logit", torch.float32)
    return b.data_ptr()


def qk_logit_bound(wq: torch.Tensor, wk: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wq, wk [L,128] bf16 per-head RMSNorm weights -> fp32 [L] bounds on |q.k|/sqrt(128) (Cauchy-Schwarz)"""
    _req(wq, "wq"), _req(wk, "wk")
    if wq.shape != wk.shape or wq.shape[-1] != 128 or not wq.is_contiguous() or not wk.is_contiguous():
        raise ValueError("wq, wk must be contiguous [L, 128]")
    L = wq.numel() // 128
    if out is None:
        out = torch.empty((L,), device=wq.device, dtype=torch.float32)
    _lib.call("drb_qk_logit_bound", wq.data_ptr(), wk.data_ptr(), out.data_ptr(), L, _stream())
    return out


def adaln_modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, out: Optional[torch.Tensor] = None,
                   add_gate: Optional[torch.Tensor] = None, add_vec: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, "x"), _req(shift, "shift"), _req(scale, "scale")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.call("drb_adaln_modulate", x.data_ptr(), out.data_ptr(), shift.data_ptr(), scale.data_ptr(), _ptr(add_gate),
              _ptr(add_vec), rows, D, _stream())
    return out


def qk_norm_rope(qkv: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, cos_tab: torch.Tensor, sin_tab: torch.Tensor,
                 num_heads: int) -> torch.Tensor:
    _req(qkv, "qkv"), _req(wq, "wq"), _req(wk, "wk"), _req(cos_tab, "cos_tab"), _req(sin_tab, "sin_tab")
    _lib.call("drb_qk_norm_rope", qkv.data_ptr(), _rows2d(qkv, "qkv"), wq.data_ptr(), wk.data_ptr(), cos_tab.data_ptr(),
              sin_tab.data_ptr(), qkv.shape[0], num_heads, _stream())
    return qkv


def gemv(w: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None,
         act: int = 0) -> torch.Tensor:
    _req(w, "w"), _req(x, "x")
    N, K = w.shape
    if out is None:
        out = torch.empty((N,), device=w.device, dtype=BF16)
    _lib.call("drb_gemv_bf16", w.data_ptr(), _rows2d(w, "w"), x.data_ptr(), out.data_ptr(), _ptr(add), N, K, act, _stream())
    return out


def gemv_batched(w: torch.Tensor, x: torch.Tensor, out: torch.Tensor, add: Optional[torch.Tensor] = None, act: int = 0,
                 n: Optional[int] = None) -> torch.Tensor:
    """w [B,N,K] contiguous; x [K] (shared) or [B,K]; out [B,N]; add [N] (shared), [B,N] or None."""
    _req(w, "w"), _req(x, "x"), _req(out, "out")
    B, N, K = w.shape
    n = N if n is None else n
    x_bs = 0 if x.ndim == 1 else x.stride(0)
    add_bs = 0 if (add is None or add.ndim == 1) else add.stride(0)
    _lib.call("drb_gemv_bf16_batched", w.data_ptr(), w.stride(1), w.stride(0), x.data_ptr(), x_bs, out.data_ptr(),
              out.stride(0), _ptr(add), add_bs, B, n, K, act, _stream())
    return out


def sigma_embedding(sigma: torch.Tensor, w_aff: torch.Tensor, e_out: torch.Tensor, emb_out: torch.Tensor) -> None:
    _req(sigma, "sigma", torch.float32), _req(w_aff, "w_aff")
    _lib.call("drb_sigma_embedding", sigma.data_ptr(), w_aff.data_ptr(), e_out.data_ptr(), emb_out.data_ptr(),
              w_aff.numel(), _stream())


def scale_patchify(x_t: torch.Tensor, sigma: torch.Tensor, tokens: torch.Tensor) -> None:
    """x_t [C,T,H,W] bf16 contiguous -> tokens[:, 0:4C] (c_in-scaled)."""
    _req(x_t, "x_t"), _req(sigma, "sigma", torch.float32), _req(tokens, "tokens")
    C, T, H, W = x_t.shape
    _lib.call("drb_scale_patchify", x_t.data_ptr(), sigma.data_ptr(), tokens.data_ptr(), _rows2d(tokens, "tokens"), C, T, H,
              W, _stream())


def patchify_condition(src: Optional[torch.Tensor], tokens: torch.Tensor, c0: int, T: int, H: int, W: int,
                       ones_channel: int = -1, zero_from: Optional[int] = None) -> None:
    _req(tokens, "tokens")
    C = 0
    if src is not None:
        _req(src, "src")
        C = src.shape[0]
    ld = _rows2d(tokens, "tokens")
    _lib.call("drb_patchify_condition", _ptr(src), tokens.data_ptr(), ld, c0, C, T, H, W, ones_channel,
              ld if zero_from is None else zero_from, _stream())


def unpatchify_euler(y_cond: torch.Tensor, y_uncond: Optional[torch.Tensor], guidance: float,
                     sigma: Optional[torch.Tensor], sigma_next: Optional[torch.Tensor], x_t: Optional[torch.Tensor],
                     x_next: Optional[torch.Tensor], f_out: Optional[torch.Tensor] = None) -> None:
    """y [S, 4C] -> F [C,T,H,W] (+ CFG) and, when x_t/x_next are given, the EDM Euler update x_t -> x_next."""
    _req(y_cond, "y_cond")
    ref = x_t if x_t is not None else f_out
    if ref is None:
        raise ValueError("pass x_t/x_next and/or f_out")
    _req(ref, "x_t/f_out")
    if not ref.is_contiguous():
        raise ValueError("latent tensors must be contiguous [C,T,H,W]")
    C, T, H, W = ref.shape
    _lib.call("drb_unpatchify_euler", y_cond.data_ptr(), _ptr(y_uncond), _rows2d(y_cond, "y_cond"), float(guidance),
              _ptr(sigma), _ptr(sigma_next), _ptr(x_t), _ptr(x_next), _ptr(f_out), C, T, H, W, _stream())


def edm_scale_input(x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    _req(x, "x"), _req(sigma, "sigma", torch.float32)
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.call("drb_edm_scale_input", x.data_ptr(), sigma.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def edm_euler_step(model_output: torch.Tensor, x: torch.Tensor, sigma: torch.Tensor, sigma_next: torch.Tensor) -> torch.Tensor:
    _req(model_output, "model_output"), _req(x, "x"), _req(sigma, "sigma", torch.float32), _req(sigma_next, "sigma_next", torch.float32)
    x, model_output = x.contiguous(), model_output.contiguous()
    out = torch.empty_like(x)
    _lib.call("drb_edm_euler_step", model_output.data_ptr(), x.data_ptr(), sigma.data_ptr(), sigma_next.data_ptr(),
              out.data_ptr(), x.numel(), _stream())
    return out


def postprocess_u8(video: torch.Tensor, normalize_normal: bool = False) -> torch.Tensor:
    """video [3,T,H,W] bf16 -> uint8 [T,H,W,3]"""
    _req(video, "video")
    _, T, H, W = video.shape
    out = torch.empty((T, H, W, 3), device=video.device, dtype=torch.uint8)
    _lib.call("drb_postprocess_u8", video.data_ptr(), out.data_ptr(), T, H, W, int(bool(normalize_normal)), _stream())
    return out


# ------------------------------------------------------------------------------------------------ tokenizer operators
def _req_cl(t: torch.Tensor, name: str) -> None:
    _req(t, name)
    if t.ndim != 4 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous channels-last activation [T, H, W, C]")


def conv3d_cl(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, *, pad_h: int = 0, pad_w: int = 0, stride_hw: int = 1,
              tmode: int = _lib.TMODE_CAUSAL, out_thw: Optional[tuple] = None, resid: Optional[torch.Tensor] = None,
              resid_mode: int = _lib.RES_NONE, stats: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              out_scale: int = 1, out_off: tuple = (0, 0)) -> torch.Tensor:
    """x [T,H,W,Cin], w [Cout,kt,kh,kw,Cin], bias [Cout] -> out [T_out, H_out*out_scale, W_out*out_scale, Cout].
    `out_thw` = positions iterated (default: same as the input); `stats` float64 [T_out,2] is accumulated into."""
    _req_cl(x, "x"), _req(w, "w"), _req(bias, "bias")
    if w.ndim != 5 or not w.is_contiguous() or w.shape[4] != x.shape[3]:
        raise ValueError(f"w must be contiguous [Cout, kt, kh, kw, Cin={x.shape[3]}], got {tuple(w.shape)}")
    T, H, W, Cin = x.shape
    Cout, kt, kh, kw, _ = w.shape
    To, Ho, Wo = out_thw if out_thw is not None else (T, H, W)
    if out is None:
        out = torch.empty((To, Ho * out_scale, Wo * out_scale, Cout), device=x.device, dtype=BF16)
    _req_cl(out, "out")
    if tuple(out.shape) != (To, Ho * out_scale, Wo * out_scale, Cout):
        raise ValueError("out has the wrong shape")
    a = _lib.Conv3dArgs()
    a.x, a.w, a.bias, a.out = x.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr()
    a.resid, a.stats = None, None
    a.resid_H = a.resid_W = 0
    if resid is not None:
        _req_cl(resid, "resid")
        if resid.shape[3] != Cout:
            raise ValueError("resid must have Cout channels")
        a.resid, a.resid_H, a.resid_W = resid.data_ptr(), resid.shape[1], resid.shape[2]
    if stats is not None:
        _req(stats, "stats", torch.float64)
        if tuple(stats.shape) != (To, 2):
            raise ValueError("stats must be float64 [T_out, 2]")
        a.stats = stats.data_ptr()
    a.T_in, a.H_in, a.W_in, a.Cin = T, H, W, Cin
    a.T_out, a.H_out, a.W_out, a.Cout = To, Ho, Wo, Cout
    a.kt, a.kh, a.kw, a.pad_h, a.pad_w, a.stride_hw, a.tmode = kt, kh, kw, pad_h, pad_w, stride_hw, tmode
    a.out_scale, a.out_off_h, a.out_off_w = out_scale, out_off[0], out_off[1]
    a.resid_mode = resid_mode
    _lib.call("drb_conv3d_cl", a, _stream())
    return out


def haar_patch(x: torch.Tensor) -> torch.Tensor:
    """x [C,T,H,W] bf16 planar -> [(T+3)//4, H//4, W//4, 64*C] channels-last"""
    _req(x, "x")
    if x.ndim != 4 or not x.is_contiguous():
        raise ValueError("x must be contiguous [C, T, H, W]")
    C, T, H, W = x.shape
    out = torch.empty(((T + 3) // 4, H // 4, W // 4, 64 * C), device=x.device, dtype=BF16)
    _lib.call("drb_haar_patch", x.data_ptr(), out.data_ptr(), C, T, H, W, _stream())
    return out


def haar_unpatch(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[Tp,Hp,Wp,64*C] channels-last -> [C, 4*Tp-3, 4*Hp, 4*Wp] planar"""
    _req_cl(x, "x")
    Tp, Hp, Wp, CC = x.shape
    if CC % 64:
        raise ValueError("channel count must be 64 * C")
    C = CC // 64
    if out is None:
        out = torch.empty((C, 4 * Tp - 3, 4 * Hp, 4 * Wp), device=x.device, dtype=BF16)
    _lib.call("drb_haar_unpatch", x.data_ptr(), out.data_ptr(), C, Tp, Hp, Wp, _stream())
    return out


def haar_unpatch_u8(x: torch.Tensor, normalize_normal: bool = False) -> torch.Tensor:
    """[Tp,Hp,Wp,192] channels-last -> uint8 [4*Tp-3, 4*Hp, 4*Wp, 3]: inverse Haar + the decode post-process in one kernel"""
    _req_cl(x, "x")
    Tp, Hp, Wp, CC = x.shape
    if CC != 192:
        raise ValueError("the fused uint8 store needs a 3-channel video (192 wavelet channels)")
    out = torch.empty((4 * Tp - 3, 4 * Hp, 4 * Wp, 3), device=x.device, dtype=torch.uint8)
    _lib.call("drb_haar_unpatch_u8", x.data_ptr(), out.data_ptr(), Tp, Hp, Wp, int(bool(normalize_normal)), _stream())
    return out


def frame_stats(x: torch.Tensor) -> torch.Tensor:
    _req_cl(x, "x")
    T = x.shape[0]
    stats = torch.empty((T, 2), device=x.device, dtype=torch.float64)
    _lib.call("drb_frame_stats_cl", x.data_ptr(), stats.data_ptr(), T, x[0].numel(), _stream())
    return stats


def groupnorm_apply(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, silu: bool,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_cl(x, "x"), _req(stats, "stats", torch.float64), _req(gamma, "gamma"), _req(beta, "beta")
    T, H, W, C = x.shape
    if gamma.numel() != C or beta.numel() != C or tuple(stats.shape) != (T, 2):
        raise ValueError("gamma/beta/stats do not match x")
    if out is None:
        out = torch.empty_like(x)
    _lib.call("drb_groupnorm_apply_cl", x.data_ptr(), out.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), T,
              H * W, C, int(bool(silu)), _stream())
    return out


def softmax_rows(s: torch.Tensor, cols: int, scale: float) -> torch.Tensor:
    _req(s, "s")
    _lib.call("drb_softmax_rows", s.data_ptr(), _rows2d(s, "s"), s.shape[0], cols, float(scale), _stream())
    return s


def transpose(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """x [rows, cols] (row-strided ok) -> out [cols, ld_out >= rows]; columns beyond `rows` are zeroed."""
    _req(x, "x"), _req(out, "out")
    _lib.call("drb_transpose_bf16", x.data_ptr(), _rows2d(x, "x"), out.data_ptr(), _rows2d(out, "out"), x.shape[0], x.shape[1],
              _stream())
    return out


def spatial_attention_d512(qkv: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv [T,H,W,1536] channels-last (q | k | v, one head of dim 512) -> softmax(q k^T / sqrt(512)) v per frame, [T,H,W,512]"""
    _req_cl(qkv, "qkv")
    T, H, W, C3 = qkv.shape
    if C3 != 1536:
        raise ValueError("the fused spatial attention is specialised for one head of dim 512 (q | k | v = 1536 channels)")
    if out is None:
        out = torch.empty((T, H, W, 512), device=qkv.device, dtype=BF16)
    _req_cl(out, "out")
    _lib.call("drb_spatial_attention_d512", qkv.data_ptr(), C3, out.data_ptr(), 512, T, H * W, _stream())
    return out


def temporal_attention(qkv: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_cl(qkv, "qkv")
    T, H, W, C3 = qkv.shape
    if out is None:
        out = torch.empty((T, H, W, C3 // 3), device=qkv.device, dtype=BF16)
    _lib.call("drb_temporal_attention_cl", qkv.data_ptr(), out.data_ptr(), T, H * W, C3 // 3, _stream())
    return out


def planar_to_cl(x: torch.Tensor, c_pad: int, scale: float = 1.0) -> torch.Tensor:
    """x [C,T,H,W] -> [T,H,W,c_pad]"""
    _req(x, "x")
    if x.ndim != 4 or not x.is_contiguous():
        raise ValueError("x must be contiguous [C, T, H, W]")
    C, T, H, W = x.shape
    out = torch.empty((T, H, W, c_pad), device=x.device, dtype=BF16)
    _lib.call("drb_planar_to_cl", x.data_ptr(), out.data_ptr(), C, c_pad, T * H * W, float(scale), _stream())
    return out


def cl_to_planar(x: torch.Tensor, C: int, scale: float = 1.0) -> torch.Tensor:
    """x [T,H,W,c_pad] -> [C,T,H,W] (first C channels)"""
    _req_cl(x, "x")
    T, H, W, c_pad = x.shape
    out = torch.empty((C, T, H, W), device=x.device, dtype=BF16)
    _lib.call("drb_cl_to_planar", x.data_ptr(), out.data_ptr(), C, c_pad, T * H * W, float(scale), _stream())
    return out


def latent_normalize(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, decode: bool) -> torch.Tensor:
    """x (..., H, W) bf16 contiguous with one (mean, std) per leading index; encode: (x - mean) / std, decode: x * std + mean"""
    _req(x, "x"), _req(mean, "mean"), _req(std, "std")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    rows = mean.numel()
    hw = x.shape[-1] * x.shape[-2]
    if std.numel() != rows or x.numel() != rows * hw or not mean.is_contiguous() or not std.is_contiguous():
        raise ValueError("one (mean, std) pair per (batch, channel, frame) row of x")
    out = torch.empty_like(x)
    _lib.call("drb_latent_normalize", x.data_ptr(), mean.data_ptr(), std.data_ptr(), out.data_ptr(), rows, hw, int(bool(decode)), _stream())
    return out


# ------------------------------------------------------------------------------------------------ context parallelism
def qk_norm_rope_scatter(qkv: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, cos_tab: torch.Tensor, sin_tab: torch.Tensor,
                         num_heads: int, dst_ptrs, dst_ld: int, row0: int) -> None:
    """norm + RoPE of the local token rows of qkv [S_loc, 3D]; head h's q/k/v go to dst_ptrs[h // (H/P)] (peer [S, dst_ld])"""
    _req(qkv, "qkv"), _req(wq, "wq"), _req(wk, "wk"), _req(cos_tab, "cos_tab"), _req(sin_tab, "sin_tab")
    if cos_tab.shape[0] != qkv.shape[0] or not cos_tab.is_contiguous() or not sin_tab.is_contiguous():
        raise ValueError("cos/sin tables must hold exactly the local token rows")
    _lib.call("drb_cp_qk_norm_rope_scatter", qkv.data_ptr(), _rows2d(qkv, "qkv"), wq.data_ptr(), wk.data_ptr(), cos_tab.data_ptr(),
              sin_tab.data_ptr(), qkv.shape[0], num_heads, _lib.ptr_array(dst_ptrs), len(dst_ptrs), dst_ld, row0, _stream())


def attention_cp(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, o_ptrs, ld_o: int, rows_per_rank: int,
                 col0: int, heads_per_batch: Optional[int] = None, batch_rows: int = 0,
                 max_abs_logit: Optional[torch.Tensor] = None, sync=None) -> None:
    """attention over the local heads and all tokens; output row r is stored to o_ptrs[r // rows_per_rank].  Batched
    sequences: `num_heads` = batch * heads_per_batch (b, h) pairs; sequence b lands at local row b * batch_rows + ..."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, n)
    ld = _rows2d(q, "q")
    if _rows2d(k, "k") != ld or _rows2d(v, "v") != ld:
        raise ValueError("q, k, v must share one row pitch")
    _lib.call("drb_attention_bf16_cp_batched", q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, _lib.ptr_array(o_ptrs), len(o_ptrs), ld_o,
              q.shape[0], k.shape[0], num_heads, rows_per_rank, col0, num_heads if heads_per_batch is None else heads_per_batch,
              batch_rows, _bound_ptr(max_abs_logit), None if sync is None else ctypes.byref(sync), _stream())


# ------------------------------------------------------------------------------------------------ environment maps (fp32)
def envmap_latlong_to_cubemap(latlong: torch.Tensor, brightness: float, flip: bool, roll_px: int, res: int = 512) -> torch.Tensor:
    """latlong [He, We, 3] fp32 -> preprocessed cube map [6, res, res, 3]"""
    _req(latlong, "latlong", torch.float32)
    if latlong.ndim != 3 or latlong.shape[2] != 3 or not latlong.is_contiguous():
        raise ValueError("latlong must be a contiguous [H, W, 3] tensor")
    cube = torch.empty((6, res, res, 3), device=latlong.device, dtype=torch.float32)
    _lib.call("drb_envmap_latlong_to_cubemap", latlong.data_ptr(), latlong.shape[0], latlong.shape[1], float(brightness), int(bool(flip)),
              int(roll_px), cube.data_ptr(), res, _stream())
    return cube


def envmap_project(cube: torch.Tensor, H: int, W: int):
    """cube [6, R, R, 3] fp32 -> (env_ldr, env_log), each [H, W, 3] in [0, 1]"""
    _req(cube, "cube", torch.float32)
    if cube.ndim != 4 or cube.shape[0] != 6 or cube.shape[1] != cube.shape[2] or cube.shape[3] != 3 or not cube.is_contiguous():
        raise ValueError("cube must be a contiguous [6, R, R, 3] tensor")
    ldr = torch.empty((H, W, 3), device=cube.device, dtype=torch.float32)
    lg = torch.empty_like(ldr)
    _lib.call("drb_envmap_project", cube.data_ptr(), cube.shape[1], ldr.data_ptr(), lg.data_ptr(), H, W, _stream())
    return ldr, lg


def envmap_tonemap(img: torch.Tensor, H: int, W: int):
    """img [Hs, Ws, 3] fp32 -> bilinear resize to (H, W) + tone mapping: (env_ldr, env_log)"""
    _req(img, "img", torch.float32)
    if img.ndim != 3 or img.shape[2] != 3 or not img.is_contiguous():
        raise ValueError("img must be a contiguous [H, W, 3] tensor")
    ldr = torch.empty((H, W, 3), device=img.device, dtype=torch.float32)
    lg = torch.empty_like(ldr)
    _lib.call("drb_envmap_tonemap", img.data_ptr(), img.shape[0], img.shape[1], ldr.data_ptr(), lg.data_ptr(), H, W, _stream())
    return ldr, lg


def qkv_gemm_norm_rope(a: torch.Tensor, w: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, cos_tab: torch.Tensor,
                       sin_tab: torch.Tensor, out: Optional[torch.Tensor] = None, peer_ptrs=None, peer_ld: int = 0,
                       row0: int = 0, batch: int = 1, sync=None) -> Optional[torch.Tensor]:
    """[q | k | v] = a[M,K] @ w[3D,K]^T with per-head RMSNorm + RoPE of q, k in the GEMM epilogue.  Plain: rows go to
    out [M, 3D].  Context-parallel (`peer_ptrs`): head h's rows go to peer_ptrs[h // (H/P)] at row row0 + s (no `out`).
    `batch` sequences of M / batch rows each are stacked along M (cos/sin tables repeated per sequence)."""
    for t, n in ((a, "a"), (w, "w"), (wq, "wq"), (wk, "wk"), (cos_tab, "cos_tab"), (sin_tab, "sin_tab")):
        _req(t, n)
    M, K = a.shape
    if w.shape[1] != K or w.shape[0] % 3:
        raise ValueError(f"shape mismatch: a {tuple(a.shape)} vs w {tuple(w.shape)}")
    D = w.shape[0] // 3
    if cos_tab.shape != (M, 128) or sin_tab.shape != (M, 128) or not cos_tab.is_contiguous() or not sin_tab.is_contiguous():
        raise ValueError("cos/sin tables must be contiguous [M, 128]")
    if peer_ptrs is None:
        if out is None:
            out = torch.empty((M, 3 * D), device=a.device, dtype=BF16)
        _req(out, "out")
        _lib.call("drb_gemm_qkv_norm_rope", a.data_ptr(), _rows2d(a, "a"), w.data_ptr(), _rows2d(w, "w"), out.data_ptr(),
                  _rows2d(out, "out"), M, D, K, wq.data_ptr(), wk.data_ptr(), cos_tab.data_ptr(), sin_tab.data_ptr(), None, 0, 0, 0,
                  _stream())
        return out
    if M % batch:
        raise ValueError("the rows do not split into `batch` equal sequences")
    _lib.call("drb_gemm_qkv_norm_rope_batched", a.data_ptr(), _rows2d(a, "a"), w.data_ptr(), _rows2d(w, "w"), None, 0, M, D, K,
              wq.data_ptr(), wk.data_ptr(), cos_tab.data_ptr(), sin_tab.data_ptr(), _lib.ptr_array(peer_ptrs), len(peer_ptrs),
              peer_ld, row0, batch, M // batch, None if sync is None else ctypes.byref(sync), _stream())
    return None


def attention_ring_block(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, state_o: torch.Tensor,
                         state_ml: torch.Tensor, first: bool, last: bool, out: Optional[torch.Tensor] = None,
                         max_abs_logit: Optional[torch.Tensor] = None) -> None:
    """one K/V block of ring attention: merges softmax(q k^T) v over this block into the fp32 running state; the last block
    writes the normalised rows to `out` [q_len, H*128]"""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, n)
    _req(state_o, "state_o", torch.float32), _req(state_ml, "state_ml", torch.float32)
    ld = _rows2d(q, "q")
    if _rows2d(k, "k") != ld or _rows2d(v, "v") != ld:
        raise ValueError("q, k, v must share one row pitch")
    if state_o.numel() != q.shape[0] * num_heads * 128 or state_ml.numel() != q.shape[0] * num_heads * 2:
        raise ValueError("state buffers do not match q_len x heads")
    if last:
        if out is None:
            raise ValueError("the last block needs `out`")
        _req(out, "out")
    _lib.call("drb_attention_bf16_ring_bounded", q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, _ptr(out) if last else None,
              _rows2d(out, "out") if last else 0, state_o.data_ptr(), state_ml.data_ptr(), q.shape[0], k.shape[0], num_heads,
              int(first), int(last), _bound_ptr(max_abs_logit), _stream())
