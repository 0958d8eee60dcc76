"""GeneralDIT denoiser of DiffusionRenderer on B200 — drop-in for the reference `CleanGeneralDIT.py`.

Same constructor kwargs, same `forward(x, timesteps, latent_condition, context_index)` contract and the same
`state_dict()` key names as the reference `CleanDiffusionRendererGeneralDIT` (CleanGeneralDIT.py:721-751), so a
reference checkpoint loads with `strict=True`.  The arithmetic is not PyTorch: every operator on the token stream is a
hand-written sm_100a kernel reached through the C ABI of libdrb200.so (include/drb200.h); this module only owns the
weights, packs them into kernel-friendly layouts, and sequences kernel launches on the current CUDA stream.

Per forward (reference call sites in brackets):
  sigma embedding + AdaLN-LoRA GEMVs for all 3L+1 sub-blocks at once   [:321-372, :483-501, :558-572]
  patchify [x | condition | ones] -> tokens, patch GEMM                  [:409-417, :669-678]
  L x { AdaLN -> fused QKV GEMM with per-head RMSNorm + RoPE in its epilogue -> flash attention -> out GEMM with gated residual;
        cross-attention collapsed to one vector per block (one-token context => softmax == 1) added inside the
        next AdaLN; AdaLN -> GEMM+GELU -> GEMM with gated residual }      [:268-306, :442-462, :492-517]
  final AdaLN -> GEMM (D -> 64) -> unpatchify                            [:567-590, :709-716]
There is no CPU or eager fallback: without libdrb200.so or off a CUDA device every call raises.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops

BF16 = torch.bfloat16


class _Weight(nn.Module):
    """A bare `.weight` parameter (the reference's bias-free nn.Linear / RMSNorm / nn.Embedding leaves)."""

    def __init__(self, *shape: int, init: str = "linear"):
        super().__init__()
        w = torch.empty(*shape)
        if w.device.type != "meta":
            if init == "ones":
                nn.init.ones_(w)
            elif init == "normal":
                nn.init.normal_(w)
            else:   # nn.Linear's default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                bound = 1.0 / (shape[-1] ** 0.5)
                nn.init.uniform_(w, -bound, bound)
        self.weight = nn.Parameter(w, requires_grad=False)


def _seq(*mods: nn.Module) -> nn.ModuleDict:
    """children named '0', '1', ... like nn.Sequential, but never called"""
    return nn.ModuleDict({str(i): m for i, m in enumerate(mods)})


class _Placeholder(nn.Module):
    """parameter-free slot (keeps the reference's Sequential numbering: e.g. adaLN_modulation.0 is nn.SiLU)"""


class _RoPE3D(nn.Module):
    """Angle table of the reference CleanRoPE3D (CleanGeneralDIT.py:86-159).

    `seq` (persistent) and the two frequency ranges are *buffers*, so `module.to(bfloat16)` casts them and the whole
    angle computation then runs in bf16 — a systematic rounding the kernels must see.  The table is
    tiny (S x 128) and depends only on the clip shape: it is built once per shape with torch ops on the device and
    consumed by drb_qk_norm_rope."""

    def __init__(self, head_dim: int):
        super().__init__()
        self.head_dim = head_dim
        self.register_buffer("seq", torch.arange(max(512, head_dim), dtype=torch.float))
        self.t_ntk_factor = 2.0

    def tables(self, T: int, H: int, W: int, dtype: torch.dtype) -> Tuple[torch.Tensor, torch.Tensor]:
        # The reference keeps the two frequency ranges as non-persistent buffers (:106-111) that follow the module's
        # dtype; they are rebuilt here in the dtype of `seq` (same values, and immune to `to_empty()` on a meta skeleton).
        dim_h = self.head_dim // 6 * 2
        dim_t = self.head_dim - 2 * dim_h
        dev, bdt = self.seq.device, self.seq.dtype
        r_s = (torch.arange(0, dim_h, 2, device=dev)[: dim_h // 2].float() / dim_h).to(bdt)
        r_t = (torch.arange(0, dim_t, 2, device=dev)[: dim_t // 2].float() / dim_t).to(bdt)
        f_s = 1.0 / (10000.0 ** r_s)
        f_t = 1.0 / ((10000.0 * self.t_ntk_factor) ** r_t)
        a_t = torch.outer(self.seq[:T], f_t)[:, None, None, :].expand(T, H, W, -1)
        a_h = torch.outer(self.seq[:H], f_s)[None, :, None, :].expand(T, H, W, -1)
        a_w = torch.outer(self.seq[:W], f_s)[None, None, :, :].expand(T, H, W, -1)
        ang = torch.cat([a_t, a_h, a_w, a_t, a_h, a_w], dim=-1).reshape(T * H * W, self.head_dim).to(dtype)
        return ang.cos().to(dtype).contiguous(), ang.sin().to(dtype).contiguous()


class CleanGeneralDIT(nn.Module):
    """Weights + launch sequencing of the GeneralDIT (reference CleanGeneralDIT.py:593-718)."""

    def __init__(self, **kwargs):
        super().__init__()
        D = self.model_channels = kwargs["model_channels"]
        L = self.num_blocks = kwargs["num_blocks"]
        Hh = self.num_heads = kwargs["num_heads"]
        self.in_channels = kwargs["in_channels"]
        self.out_channels = kwargs["out_channels"]
        Cc = self.crossattn_emb_channels = kwargs["crossattn_emb_channels"]
        if kwargs.get("block_config", "FA-CA-MLP").upper().replace("_", "-") != "FA-CA-MLP":
            raise ValueError("only block_config 'FA-CA-MLP' is implemented (the only one the reference configs use)")
        self.mlp_ratio = kwargs.get("mlp_ratio", 4.0)
        self.additional_concat_ch = kwargs.get("additional_concat_ch", 0)
        self.concat_padding_mask = kwargs.get("concat_padding_mask", True)
        self.patch_spatial = kwargs["patch_spatial"]
        self.patch_temporal = kwargs["patch_temporal"]
        R = self.adaln_lora_dim = kwargs.get("adaln_lora_dim", 256)
        if self.patch_spatial != 2 or self.patch_temporal != 1:
            raise ValueError("only patch_spatial=2, patch_temporal=1 is implemented (reference configs)")
        if D % Hh != 0 or D // Hh != 128:
            raise ValueError("head_dim must be 128 (the attention kernel is specialised for it)")
        if D % 256 != 0 or D > 4096:
            raise ValueError("model_channels must be a multiple of 256, at most 4096")
        if not kwargs.get("use_adaln_lora", True) or not kwargs.get("affline_emb_norm", True):
            raise ValueError("only the AdaLN-LoRA + affine-norm variant is implemented (DiffusionRenderer)")
        self.head_dim = 128
        self.hidden = int(D * self.mlp_ratio)
        self.in_total = self.in_channels + self.additional_concat_ch + (1 if self.concat_padding_mask else 0)
        self.patch_dim = self.in_total * 4
        self.out_patch_dim = self.out_channels * 4

        # ---- parameter tree with the reference's key names --------------------------------------------------
        self.x_embedder = nn.ModuleDict({"proj": nn.ModuleDict({"1": _Weight(D, self.patch_dim)})})
        te = nn.Module()
        te.linear_1 = _Weight(D, D)
        te.linear_2 = _Weight(3 * D, D)
        self.t_embedder = nn.ModuleDict({"1": te})
        self.pos_embedder = _RoPE3D(self.head_dim)

        def attn(ctx_dim: int) -> nn.Module:
            a = nn.Module()
            a.to_q = _seq(_Weight(D, D), _Weight(self.head_dim, init="ones"))
            a.to_k = _seq(_Weight(D, ctx_dim), _Weight(self.head_dim, init="ones"))
            a.to_v = _seq(_Weight(D, ctx_dim))
            a.to_out = _seq(_Weight(D, D))
            holder = nn.Module()
            holder.attn = a
            return holder

        def sub_block(inner: nn.Module, n_mod: int = 3) -> nn.Module:
            sb = nn.Module()
            sb.block = inner
            sb.adaLN_modulation = nn.ModuleDict({"1": _Weight(R, D), "2": _Weight(n_mod * D, R)})
            return sb

        blocks = {}
        for i in range(L):
            mlp = nn.Module()
            mlp.layer1 = _Weight(self.hidden, D)
            mlp.layer2 = _Weight(D, self.hidden)
            blk = nn.Module()
            blk.blocks = nn.ModuleList([sub_block(attn(D)), sub_block(attn(Cc)), sub_block(mlp)])
            blocks[f"block{i}"] = blk
        self.blocks = nn.ModuleDict(blocks)
        fl = nn.Module()
        fl.linear = _Weight(self.out_patch_dim, D)
        fl.adaLN_modulation = nn.ModuleDict({"1": _Weight(R, D), "2": _Weight(2 * D, R)})
        self.final_layer = fl
        self.affline_norm = _Weight(D, init="ones")

        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self._cp = None     # context_parallel.ContextParallel when one video is split over several GPUs
        self.fuse_qkv_epilogue = True   # False: QKV GEMM, then the stand-alone norm + RoPE (+ scatter) kernel (A/B measurements)

    @torch.no_grad()
    def init_weights_(self, seed: int = 0) -> "CleanGeneralDIT":
        """Deterministic random init IN PLACE on whatever device the parameters live on (used for the random-init
        BASELINE configs: a 7B model is materialised with `to_empty(device='cuda')` and filled here, never on the
        host).  Linear-like weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); norm weights: 1 + 0.1 N(0,1); embedding N(0,1)."""
        g = torch.Generator(device=self.affline_norm.weight.device)
        for i, (name, p) in enumerate(self.named_parameters()):
            g.manual_seed(seed * 100003 + i)
            if p.ndim == 1:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32))
            elif name.startswith("context_embedding"):
                p.copy_(torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32))
            else:
                bound = 1.0 / (p.shape[-1] ** 0.5)
                for r0 in range(0, p.shape[0], 4096):      # chunked: bounds the fp32 staging buffer
                    blk = p[r0:r0 + 4096]
                    blk.copy_((torch.rand(blk.shape, generator=g, device=p.device, dtype=torch.float32) * 2 - 1) * bound)
        self.pos_embedder.seq.copy_(torch.arange(self.pos_embedder.seq.numel(), device=self.pos_embedder.seq.device))
        return self

    # ------------------------------------------------------------------------------------------------ packing
    def _sub(self, i: int, j: int) -> nn.Module:
        return self.blocks[f"block{i}"].blocks[j]

    def _pack_key(self):
        w = self.affline_norm.weight
        return (w.device, w.dtype, w.data_ptr(), self._sub(0, 0).block.attn.to_q["0"].weight.data_ptr())

    def _ensure_packed(self) -> Dict[str, torch.Tensor]:
        """Build the kernel-side weight layout once per (device, dtype) and re-point the module's parameters at views
        of it, so `state_dict()` / `load_state_dict()` keep working and no second copy of the 7B weights exists.
          qkv  [L, 3D, D]   to_q | to_k | to_v rows           (one GEMM instead of three, CleanGeneralDIT.py:273-276)
          wo   [L, D, D], w1 [L, 4D, D], w2 [L, D, 4D]
          ca_v [L, D, Cc], ca_o [L, D, D]                     (cross-attention collapses to to_out(to_v(ctx)))
          mod_a [3L+1, R, D], mod_b [3L+1, 3D, R]             (all AdaLN-LoRA pairs; the final layer's 2D rows zero-padded)
          wx   [D, patch_dim rounded up to 8]                 (TMA needs 16-byte row pitch: 132 -> 136, 612 -> 616)
        """
        key = self._pack_key()
        if self._packed is not None and self._packed_key == key:
            return self._packed
        dev, dt = key[0], key[1]
        if dev.type != "cuda":
            raise RuntimeError("the B200 GeneralDIT runs on a CUDA device only (no CPU fallback); call .to('cuda') first")
        if dt != BF16:
            raise RuntimeError(f"the B200 GeneralDIT computes in bfloat16; got parameters in {dt} — call .to(torch.bfloat16)")
        if not _lib.load().drb_device_supported():
            raise RuntimeError("libdrb200.so targets sm_100a (B200) only; the current CUDA device is not compute capability 10.x")
        D, L, R, Cc, Hd = self.model_channels, self.num_blocks, self.adaln_lora_dim, self.crossattn_emb_channels, self.hidden
        new = lambda *s: torch.zeros(*s, device=dev, dtype=dt)
        P = {"qkv": new(L, 3 * D, D), "wo": new(L, D, D), "w1": new(L, Hd, D), "w2": new(L, D, Hd),
             "ca_v": new(L, D, Cc), "ca_o": new(L, D, D), "mod_a": new(3 * L + 1, R, D), "mod_b": new(3 * L + 1, 3 * D, R),
             "qn": new(L, 128), "kn": new(L, 128)}

        def adopt(param: nn.Parameter, view: torch.Tensor):
            view.copy_(param.data)
            param.data = view

        for i in range(L):
            sa, ca, mlp = self._sub(i, 0), self._sub(i, 1), self._sub(i, 2)
            adopt(sa.block.attn.to_q["0"].weight, P["qkv"][i, :D])
            adopt(sa.block.attn.to_k["0"].weight, P["qkv"][i, D:2 * D])
            adopt(sa.block.attn.to_v["0"].weight, P["qkv"][i, 2 * D:])
            adopt(sa.block.attn.to_out["0"].weight, P["wo"][i])
            adopt(sa.block.attn.to_q["1"].weight, P["qn"][i])
            adopt(sa.block.attn.to_k["1"].weight, P["kn"][i])
            adopt(ca.block.attn.to_v["0"].weight, P["ca_v"][i])
            adopt(ca.block.attn.to_out["0"].weight, P["ca_o"][i])
            adopt(mlp.block.layer1.weight, P["w1"][i])
            adopt(mlp.block.layer2.weight, P["w2"][i])
            for j, sb in enumerate((sa, ca, mlp)):
                adopt(sb.adaLN_modulation["1"].weight, P["mod_a"][3 * i + j])
                adopt(sb.adaLN_modulation["2"].weight, P["mod_b"][3 * i + j])
        adopt(self.final_layer.adaLN_modulation["1"].weight, P["mod_a"][3 * L])
        adopt(self.final_layer.adaLN_modulation["2"].weight, P["mod_b"][3 * L, :2 * D])
        kpad = (self.patch_dim + 7) // 8 * 8
        P["wx"] = new(D, kpad)
        adopt(self.x_embedder["proj"]["1"].weight, P["wx"][:, :self.patch_dim])
        self._packed, self._packed_key = P, self._pack_key()
        self._ws.clear()
        return P

    def enable_context_parallel(self, cp) -> None:
        """Split every forward's token sequence over the ranks of `cp` (context_parallel.ContextParallel, or None to go
        back to one GPU).  All ranks then call forward / the sampler with identical arguments and get identical results."""
        if cp is not None and self.num_heads % cp.world:
            raise ValueError(f"{self.num_heads} heads do not split over {cp.world} ranks")
        self._cp = cp
        self._ws.clear()

    def _workspace(self, T: int, H: int, W: int, dev, cp=None, batch: int = 1) -> Dict[str, torch.Tensor]:
        """Activation buffers for one clip shape (allocated once, reused every forward).  `batch` independent token
        sequences (the five G-buffer passes of a clip, cond + uncond under CFG) are stacked along the row axis: every
        token buffer holds batch * S rows, sequence b in rows [b*S, (b+1)*S).  With `cp`, T counts the latent frames of
        THIS rank: S is the local token count, `a2a` [S*P, 3*batch*D/P] receives this rank's heads of every token of
        every sequence, and `attn` / `a2a` are peer-mapped so that the other ranks' kernels store into them directly."""
        key = (T, H, W, dev, id(cp), batch)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        D, L = self.model_channels, self.num_blocks
        S = T * (H // 2) * (W // 2)
        R = batch * S
        new = lambda *s, dtype=BF16: torch.empty(*s, device=dev, dtype=dtype)
        kpad = self._packed["wx"].shape[1]
        world, rank = (cp.world, cp.rank) if cp is not None else (1, 0)
        if batch > 1 and cp is not None and cp.mode != "ring" and S < 32:
            raise ValueError("batched sequences under context parallelism need >= 32 local tokens per sequence")
        cos, sin = self.pos_embedder.tables(T * world, H // 2, W // 2, BF16)
        cos, sin = cos[rank * S:(rank + 1) * S].repeat(batch, 1).contiguous(), sin[rank * S:(rank + 1) * S].repeat(batch, 1).contiguous()
        ws = {
            "S": S, "B": batch, "tok": torch.zeros(R, kpad, device=dev, dtype=BF16), "x": new(R, D), "xm": new(R, D),
            "h": new(R, self.hidden), "y": new(R, self.out_patch_dim), "cos": cos, "sin": sin,
            "e": new(D), "emb": new(D), "t1": new(D), "lora": new(3 * D), "mod_h": new(3 * L + 1, self.adaln_lora_dim),
            "mod": new(3 * L + 1, 3 * D), "ca_tmp": new(batch, L, D), "ca_vec": new(batch, L, D),
            "sigma": new(1, dtype=torch.float32), "qk_bound": new(L, dtype=torch.float32), "cp": cp,
        }
        if cp is None:
            ws["attn"], ws["qkv"] = new(R, D), new(R, 3 * D)
        elif cp.mode == "ring":
            # every rank's [q | k | v] rows are visible to the others; remote K/V blocks land in two staging buffers of the
            # same row pitch (only their k | v columns are written); fp32 running softmax state of the local query rows
            ws["attn"] = new(R, D)
            ws["qkv"], ws["qkv_peers"] = cp.alloc_views("qkv", (R, 3 * D))
            ws["kv_stage"] = [new(S, 3 * D), new(S, 3 * D)]
            ws["ring_o"] = new(S, D, dtype=torch.float32)
            ws["ring_ml"] = new(S, self.num_heads, 2, dtype=torch.float32)
        else:
            ws["attn"], ws["attn_ptrs"] = cp.alloc("attn", (R, D))
            ws["a2a"], ws["a2a_ptrs"] = cp.alloc("a2a", (S * world, 3 * batch * D // world))
        self._ws = {key: ws}   # one live shape at a time: the buffers are large (MLP hidden = 0.92 GB at 57x704x1280)
        return ws

    # ------------------------------------------------------------------------------------------------ stages
    def prepare_condition(self, ws, latent_condition: Optional[torch.Tensor], T: int, H: int, W: int, b: int = 0) -> None:
        """Constant token features of sequence `b`: condition channels, ones padding mask, zero K-padding (:669-675)."""
        c0 = self.in_channels
        ones = c0 + self.additional_concat_ch if self.concat_padding_mask else -1
        cond = None
        if latent_condition is not None and self.additional_concat_ch > 0:
            cond = latent_condition.reshape(-1, T, H, W)
            if cond.shape[0] != self.additional_concat_ch:
                raise ValueError(f"latent_condition has {cond.shape[0]} channels, the net expects {self.additional_concat_ch}")
            cond = cond.to(BF16).contiguous()
        ops.patchify_condition(cond, self._rows(ws, "tok", b), c0, T, H, W, ones_channel=ones, zero_from=self.patch_dim)

    @staticmethod
    def _rows(ws, name: str, b: int) -> torch.Tensor:
        """rows of sequence b in a token buffer"""
        S = ws["S"]
        return ws[name][b * S:(b + 1) * S]

    def prepare_context(self, ws, ctx: Optional[torch.Tensor], b: int = 0) -> bool:
        """ca_vec[b, i] = to_out_i(to_v_i(ctx)) for every block (reference :268-306 with one key: softmax == 1).
        Returns False when the context is all zeros (forward renderer): the sub-block is then the identity."""
        if ctx is None:
            return False
        P = self._packed
        ops.gemv_batched(P["ca_v"], ctx.reshape(-1).contiguous(), ws["ca_tmp"][b])
        ops.gemv_batched(P["ca_o"], ws["ca_tmp"][b], ws["ca_vec"][b])
        return True

    def modulation(self, ws, sigma_dev: torch.Tensor) -> None:
        """Everything that depends only on sigma: embedding, AdaLN-LoRA vector, all 3L+1 (shift, scale, gate) rows —
        shared by every sequence of the batch — and the per-layer certificate of the max-free softmax (from the q/k norm
        weights as they are NOW, so that re-loaded weights can never run under a stale bound)."""
        P = self._packed
        ops.sigma_embedding(sigma_dev, self.affline_norm.weight, ws["e"], ws["emb"])
        ops.gemv(self.t_embedder["1"].linear_1.weight, ws["e"], out=ws["t1"])
        ops.gemv(self.t_embedder["1"].linear_2.weight, ws["t1"], out=ws["lora"], act=1)
        ops.gemv_batched(P["mod_a"], ws["emb"], ws["mod_h"], act=1)
        ops.gemv_batched(P["mod_b"], ws["mod_h"], ws["mod"], add=ws["lora"])
        ops.qk_logit_bound(P["qn"], P["kn"], out=ws["qk_bound"])

    def stage_embed(self, ws) -> None:
        ops.gemm(ws["tok"], self._packed["wx"], out=ws["x"])

    def stage_pre_attention(self, ws, i: int, sync=None) -> None:
        """AdaLN -> fused QKV GEMM whose epilogue does the per-head RMSNorm + RoPE of q and k (under context parallelism
        it also stores every head's rows straight into the GPU that owns the head: the all-to-all is the epilogue)"""
        P, D, cp, B = self._packed, self.model_channels, ws["cp"], ws["B"]
        m_sa = ws["mod"][3 * i]
        ops.adaln_modulate(ws["x"], m_sa[:D], m_sa[D:2 * D], out=ws["xm"])
        if not self.fuse_qkv_epilogue:
            if B > 1 and cp is not None and cp.mode != "ring":
                raise ValueError("the stand-alone scatter kernel (A/B reference) handles one sequence; use fuse_qkv_epilogue")
            qkv = ws.get("qkv")
            if qkv is None:
                qkv = ws["qkv"] = torch.empty(ws["S"], 3 * D, device=ws["x"].device, dtype=BF16)
            ops.gemm(ws["xm"], P["qkv"][i], out=qkv)
            if cp is None or cp.mode == "ring":
                ops.qk_norm_rope(qkv, P["qn"][i], P["kn"][i], ws["cos"], ws["sin"], self.num_heads)
            else:
                ops.qk_norm_rope_scatter(qkv, P["qn"][i], P["kn"][i], ws["cos"], ws["sin"], self.num_heads, ws["a2a_ptrs"],
                                         3 * D // cp.world, cp.rank * ws["S"])
        elif cp is None or cp.mode == "ring":
            ops.qkv_gemm_norm_rope(ws["xm"], P["qkv"][i], P["qn"][i], P["kn"][i], ws["cos"], ws["sin"], out=ws["qkv"])
        else:
            ops.qkv_gemm_norm_rope(ws["xm"], P["qkv"][i], P["qn"][i], P["kn"][i], ws["cos"], ws["sin"], peer_ptrs=ws["a2a_ptrs"],
                                   peer_ld=3 * B * D // cp.world, row0=cp.rank * ws["S"], batch=B, sync=sync)

    def stage_attention(self, ws, i: int, timers: Optional[list] = None, sync=None) -> None:
        D, Hh, cp, B, S = self.model_channels, self.num_heads, ws["cp"], ws["B"], ws["S"]
        bound = ws["qk_bound"][i:i + 1]
        if timers is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if cp is None:
            for b in range(B):
                qkv = ws["qkv"][b * S:(b + 1) * S]
                ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], Hh, out=ws["attn"][b * S:(b + 1) * S], max_abs_logit=bound)
        elif cp.mode == "ring":
            for b in range(B):          # the ring schedule runs per sequence (staging buffers and running state are reused)
                self._ring_attention(ws, b, bound)
        else:
            # (sequence, local head) pairs are the attention problems of this rank: B * H/P "heads" over all S tokens
            Hp, a2a = Hh // cp.world, ws["a2a"]
            w = B * Hp * 128
            ops.attention_cp(a2a[:, :w], a2a[:, w:2 * w], a2a[:, 2 * w:], B * Hp, ws["attn_ptrs"], D, S, cp.rank * Hp * 128,
                             heads_per_batch=Hp, batch_rows=S, max_abs_logit=bound, sync=sync)
        if timers is not None:
            ev[1].record()
            timers.append(ev)

    def _ring_attention(self, ws, b: int = 0, bound: Optional[torch.Tensor] = None) -> None:
        """Ring schedule over peer memory for sequence b: block s of rank r is the K/V of rank (r - s) mod P.  While block s
        is attended to on the current stream, block s + 1 is pulled from its owner's [q | k | v] buffer (k | v columns only)
        into the other staging buffer on the copy stream; the attention kernel's ring epilogue merges the blocks."""
        D, Hh, cp, S = self.model_channels, self.num_heads, ws["cp"], ws["S"]
        P, r = cp.world, cp.rank
        rows = slice(b * S, (b + 1) * S)
        main, side = torch.cuda.current_stream(), cp.copy_stream
        q = ws["qkv"][rows, :D]
        ready = torch.cuda.Event()
        ready.record(main)                    # the barrier before this stage: every rank's K/V rows are in place
        side.wait_event(ready)
        copied, freed = {}, {}
        for s in range(P):
            if s + 1 < P:                     # prefetch the next block
                nxt = ws["kv_stage"][(s + 1) & 1]
                if (s - 1) in freed:          # the attention that read this staging buffer two steps ago must be done
                    side.wait_event(freed[s - 1])
                with torch.cuda.stream(side):
                    nxt[:, D:].copy_(ws["qkv_peers"][(r - s - 1) % P][rows, D:], non_blocking=True)
                    copied[s + 1] = torch.cuda.Event()
                    copied[s + 1].record(side)
            kv = ws["qkv"][rows] if s == 0 else ws["kv_stage"][s & 1]
            if s > 0:
                main.wait_event(copied[s])
            ops.attention_ring_block(q, kv[:, D:2 * D], kv[:, 2 * D:], Hh, ws["ring_o"], ws["ring_ml"], first=(s == 0),
                                     last=(s == P - 1), out=ws["attn"][rows], max_abs_logit=bound)
            freed[s] = torch.cuda.Event()
            freed[s].record(main)

    def stage_post_attention(self, ws, i: int, use_ca: bool, sync=None) -> None:
        """out-projection with gated residual; cross-attention vector + AdaLN; MLP with gated residual"""
        P, D, B, S = self._packed, self.model_channels, ws["B"], ws["S"]
        x, xm, h, mod = ws["x"], ws["xm"], ws["h"], ws["mod"]
        m_sa, m_ca, m_mlp = mod[3 * i], mod[3 * i + 1], mod[3 * i + 2]
        ops.gemm(ws["attn"], P["wo"][i], out=x, epilogue=_lib.EPI_GATED_RESIDUAL, resid=x, gate=m_sa[2 * D:], sync=sync)
        if use_ca:      # the cross-attention vector is the only per-sequence term of a block
            for b in range(B):
                ops.adaln_modulate(x[b * S:(b + 1) * S], m_mlp[:D], m_mlp[D:2 * D], out=xm[b * S:(b + 1) * S],
                                   add_gate=m_ca[2 * D:], add_vec=ws["ca_vec"][b, i])
        else:
            ops.adaln_modulate(x, m_mlp[:D], m_mlp[D:2 * D], out=xm)
        ops.gemm(xm, P["w1"][i], out=h, epilogue=_lib.EPI_GELU)
        ops.gemm(h, P["w2"][i], out=x, epilogue=_lib.EPI_GATED_RESIDUAL, resid=x, gate=m_mlp[2 * D:])

    def stage_final(self, ws) -> torch.Tensor:
        D = self.model_channels
        m_f = ws["mod"][3 * self.num_blocks]
        ops.adaln_modulate(ws["x"], m_f[:D], m_f[D:2 * D], out=ws["xm"])
        ops.gemm(ws["xm"], self.final_layer.linear.weight, out=ws["y"])
        return ws["y"]

    def run_blocks(self, ws, use_ca: bool, timers: Optional[list] = None) -> torch.Tensor:
        """tokens -> y [B*S, 64].  Consumes ws['tok'] / ws['mod'] / ws['ca_vec'].
        `timers` (bench only): a list that receives one (start, end) CUDA-event pair per attention stage.
        Under context parallelism the P2P stores of the fused exchange are ordered twice per block: q / k / v rows must have
        landed before the attention reads them, its output rows before the out-projection.  With `cp.fused_sync` that is a
        flag published by the producer kernel's last CTA and awaited by the consumer kernel's TMA-producer thread (no launch
        in between); otherwise a stand-alone device barrier kernel."""
        cp = ws["cp"]
        fused = cp is not None and getattr(cp, "fused_sync", False) and cp.mode == "ulysses" and self.fuse_qkv_epilogue
        self.stage_embed(ws)
        for i in range(self.num_blocks):
            if fused:
                e_q, e_a = cp.next_epoch(_lib.CP_SLOT_QKV), cp.next_epoch(_lib.CP_SLOT_ATTN)
                self.stage_pre_attention(ws, i, sync=cp.sync(signal=(_lib.CP_SLOT_QKV, e_q)))
                self.stage_attention(ws, i, timers, sync=cp.sync(wait=(_lib.CP_SLOT_QKV, e_q), signal=(_lib.CP_SLOT_ATTN, e_a)))
                self.stage_post_attention(ws, i, use_ca, sync=cp.sync(wait=(_lib.CP_SLOT_ATTN, e_a)))
                continue
            self.stage_pre_attention(ws, i)
            if cp is not None:
                cp.barrier()
            self.stage_attention(ws, i, timers)
            if cp is not None:
                cp.barrier()
            self.stage_post_attention(ws, i, use_ca)
        return self.stage_final(ws)

    def denoise_step(self, ws, x: torch.Tensor, sigma: torch.Tensor, sigma_next: torch.Tensor, use_ca: bool,
                     timers: Optional[list] = None, guidance: float = 0.0) -> None:
        """One EDM Euler step in place on x [16,T,H,W] or [N,16,T,H,W] (model_diffusion_renderer.py:224-234): sigma-only
        vectors, c_in scale + patchify, ONE transformer pass over all sequences of the workspace, unpatchify (+ CFG) +
        Euler.  Guidance 0: the workspace holds N sequences (one per latent).  Guidance > 0: 2N sequences, the N
        conditional ones first, each latent feeding its cond and its uncond sequence (:230-232 as one batched forward).
        sigma / sigma_next: fp32 device scalars."""
        xs = x if x.ndim == 5 else x.unsqueeze(0)
        N, B, S = xs.shape[0], ws["B"], ws["S"]
        if B != (2 * N if guidance > 0 else N):
            raise ValueError(f"workspace holds {B} sequences, the step needs {2 * N if guidance > 0 else N}")
        self.modulation(ws, sigma)
        for b in range(B):
            ops.scale_patchify(xs[b % N], sigma, self._rows(ws, "tok", b))
        y = self.run_blocks(ws, use_ca, timers)
        for n in range(N):
            y_c = y[n * S:(n + 1) * S]
            y_u = y[(N + n) * S:(N + n + 1) * S] if guidance > 0 else None
            ops.unpatchify_euler(y_c, y_u, guidance, sigma, sigma_next, xs[n], xs[n])

    def _check_input(self, x: torch.Tensor) -> Tuple[int, int, int]:
        if x.ndim != 5:
            raise ValueError(f"expected a 5-D latent (B, C, T, H, W), got {tuple(x.shape)}")
        B, C, T, H, W = x.shape
        if B != 1:
            raise ValueError("batch size must be 1 (the reference sampler is batch-1 as well: model_diffusion_renderer.py:222)")
        if C != self.in_channels:
            raise ValueError(f"expected {self.in_channels} latent channels, got {C}")
        if H % 2 or W % 2:
            raise ValueError("latent height and width must be even (2x2 patches)")
        return T, H, W

    def forward(self, x, timesteps, crossattn_emb, latent_condition, **kwargs):
        """reference CleanGeneralDIT.forward (:656-718): x is the c_in-scaled noisy latent; returns F (B,16,T,H,W).
        With context parallelism enabled every rank passes the same full tensors, computes its frame slice and
        receives the full F."""
        T, H, W = self._check_input(x)
        self._ensure_packed()
        cp = self._cp
        t0, t1 = (0, T)
        if cp is not None:
            from .context_parallel import shard_frames
            t0, t1 = shard_frames(T, cp.rank, cp.world)
            x = x[:, :, t0:t1]
            if latent_condition is not None:
                latent_condition = latent_condition[:, :, t0:t1]
        Tl = t1 - t0
        ws = self._workspace(Tl, H, W, x.device, cp)
        ws["sigma"].copy_(torch.as_tensor(timesteps, dtype=torch.float32).reshape(-1)[:1])
        self.modulation(ws, ws["sigma"])
        self.prepare_condition(ws, latent_condition, Tl, H, W)
        ops.patchify_condition(x.reshape(-1, Tl, H, W).to(BF16).contiguous(), ws["tok"], 0, Tl, H, W)
        use_ca = self.prepare_context(ws, crossattn_emb)
        y = self.run_blocks(ws, use_ca)
        out = torch.empty((self.out_channels, Tl, H, W), device=x.device, dtype=BF16)
        ops.unpatchify_euler(y, None, 0.0, None, None, None, None, f_out=out)
        if cp is not None:
            out = cp.all_gather_frames(out)
        return out.unsqueeze(0)


class CleanDiffusionRendererGeneralDIT(CleanGeneralDIT):
    """reference CleanGeneralDIT.py:721-751 — context_index -> one context token (or zeros for the forward renderer)."""

    def __init__(self, additional_concat_ch: int = 16, use_context_embedding: bool = True, **kwargs):
        kwargs["use_adaln_lora"] = True
        kwargs["adaln_lora_dim"] = 256
        super().__init__(additional_concat_ch=additional_concat_ch, **kwargs)
        self.use_context_embedding = use_context_embedding
        if use_context_embedding:
            self.context_embedding = _Weight(16, kwargs["crossattn_emb_channels"], init="normal")

    def context_token(self, context_index) -> Optional[torch.Tensor]:
        """(1024,) context vector of this pass, or None when the net has no context embedding (zeros in the reference)."""
        if not self.use_context_embedding:
            return None
        if context_index is None:
            raise TypeError("forward() missing required argument 'context_index' (inverse renderer)")
        # device-side lookup (no .item(): the public forward must not synchronise with the host); the reference casts
        # bf16 -> long (:736)
        w = self.context_embedding.weight
        idx = torch.as_tensor(context_index).reshape(-1)[:1].to(device=w.device, non_blocking=True).long()
        return w.index_select(0, idx)[0]

    def forward(self, x, timesteps, latent_condition, context_index=None, **kwargs):
        return super().forward(x=x, timesteps=timesteps, crossattn_emb=self.context_token(context_index),
                               latent_condition=latent_condition, **kwargs)
