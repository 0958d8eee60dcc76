"""CV8x8x8 causal video tokenizer of DiffusionRenderer on B200 — drop-in for the reference `CleanVAE.py`.

The reference class is a pass-through wrapper of `diffusers.AutoencoderKLCosmos` (`CleanVAE.py:9-67`): `encode(x)
.latent_dist.sample()` and `decode(z).sample` on 5-D (B, C, T, H, W) tensors.  Here `AutoencoderKLCosmos` is this
module's own class: it holds the parameters under the diffusers `state_dict()` key names, packs
them once into the kernels' layouts, and sequences sm_100a kernels through the C ABI of libdrb200.so:

  Haar patch / unpatch                     drb_haar_patch / drb_haar_unpatch
  every CosmosCausalConv3d                 drb_conv3d_cl  (implicit GEMM on tcgen05; bias, skip term, avg-pool term and
                                                           the next GroupNorm's statistics fused in the epilogue)
  CosmosCausalGroupNorm (+ SiLU)           statistics from the producing convolution, one apply pass
  mid-block spatial attention (dim 512)    drb_gemm_bf16 -> drb_softmax_rows -> drb_gemm_bf16 per frame
  mid-block causal temporal attention      drb_temporal_attention_cl
  nearest-x2 upsample + 3x3 convolution    four sub-pixel 2x2 convolutions on the source resolution (pre-summed taps)

Activations are channels-last bf16 [T, H, W, C]; tensors narrower than 64 channels (the 16-channel latent) are
zero-padded to 64 so that every convolution sees whole 64-channel TMA boxes.  There is no CPU / eager fallback.
"""
from __future__ import annotations

import json
import math
import os
from types import SimpleNamespace
from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops

BF16 = torch.bfloat16

DEFAULT_CONFIG = dict(            # reference VAE_config.json (the values that define the network)
    in_channels=3, out_channels=3, latent_channels=16,
    encoder_block_out_channels=(128, 256, 512, 512), decode_block_out_channels=(256, 512, 512, 512),
    attention_resolutions=(32,), resolution=1024, num_layers=2, patch_size=4, patch_type="haar",
    scaling_factor=1.0, spatial_compression_ratio=8, temporal_compression_ratio=8)


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


# ------------------------------------------------------------------------------------------------ topology
def _block_plan(channels, n_sp: int, n_tp: int, decoder: bool) -> List[dict]:
    """channels and resampling flags of the down / up blocks (CosmosEncoder3d / CosmosDecoder3d __init__)"""
    plan = []
    for i in range(len(channels) - 1):
        sp = tp = False
        if i < len(channels) - 2:
            sp, tp = ((0 < i < n_sp + 1), (0 < i < n_tp + 1)) if decoder else (i < n_sp, i < n_tp)
        plan.append(dict(cin=channels[i], cout=channels[i + 1], spatial=sp, temporal=tp))
    return plan


def _convproj_shapes(p: str, cin: int, cout: int):
    yield f"{p}.conv_s.weight", (cout, cin, 1, 3, 3)
    yield f"{p}.conv_s.bias", (cout,)
    yield f"{p}.conv_t.weight", (cout, cout, 3, 1, 1)
    yield f"{p}.conv_t.bias", (cout,)


def _resnet_shapes(p: str, cin: int, cout: int):
    yield f"{p}.norm1.norm.weight", (cin,)
    yield f"{p}.norm1.norm.bias", (cin,)
    yield from _convproj_shapes(f"{p}.conv1", cin, cout)
    yield f"{p}.norm2.norm.weight", (cout,)
    yield f"{p}.norm2.norm.bias", (cout,)
    yield from _convproj_shapes(f"{p}.conv2", cout, cout)
    if cin != cout:
        yield f"{p}.conv_shortcut.weight", (cout, cin, 1, 1, 1)
        yield f"{p}.conv_shortcut.bias", (cout,)


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
    # Downsample: conv1 spatial (1,3,3)/s2, conv2 temporal (3,1,1)/s2.  Upsample: conv1 temporal, conv2 spatial.  conv3 1x1x1.
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


class _Node(nn.Module):
    """anonymous container: the parameter tree only has to reproduce the diffusers key names"""


class AutoencoderKLCosmos(nn.Module):
    """Parameters + launch sequencing of the Cosmos causal video tokenizer (diffusers AutoencoderKLCosmos restated)."""

    def __init__(self, **config):
        super().__init__()
        cfg = dict(DEFAULT_CONFIG)
        cfg.update({k: v for k, v in config.items() if k in DEFAULT_CONFIG})
        # per-(channel, latent frame) statistics of a 16-latent-frame chunk (VAE_config.json:21-536): not part of the network,
        # used only by the opt-in per-chunk normalisation of CleanVAE.encode_chunked / decode_chunked
        cfg["latents_mean"] = list(config["latents_mean"]) if config.get("latents_mean") is not None else None
        cfg["latents_std"] = list(config["latents_std"]) if config.get("latents_std") is not None else None
        if cfg["patch_type"] != "haar" or cfg["patch_size"] != 4:
            raise ValueError("only the Haar patcher with patch_size 4 is implemented (CV8x8x8)")
        self.config = SimpleNamespace(**cfg)
        c = self.config
        n_sp = int(math.log2(c.spatial_compression_ratio)) - int(math.log2(c.patch_size))
        n_tp = int(math.log2(c.temporal_compression_ratio)) - int(math.log2(c.patch_size))
        self.enc_plan = _block_plan(list(c.encoder_block_out_channels), n_sp, n_tp, decoder=False)
        self.dec_plan = _block_plan(list(reversed(c.decode_block_out_channels)), n_sp, n_tp, decoder=True)
        self.inner_dim = c.in_channels * c.patch_size ** 3
        for key, shape in self.param_shapes():
            self._add_param(key, shape)
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None
        self.fused_spatial_attention = True    # False: the GEMM -> softmax -> GEMM form also at dim 512 (A/B measurements)

    # ---------------------------------------------------------------------------------------------- parameters
    def param_shapes(self) -> Iterator[Tuple[str, Tuple[int, ...]]]:
        """(key, shape) of every state_dict entry, module order.  The resolutions visited are 256, 128, 64 (never 32 =
        attention_resolutions), so only the mid blocks carry attention."""
        c = self.config
        ech = list(c.encoder_block_out_channels)
        yield from _convproj_shapes("encoder.conv_in", self.inner_dim, ech[0])
        for i, b in enumerate(self.enc_plan):
            for j in range(c.num_layers):
                yield from _resnet_shapes(f"encoder.down_blocks.{i}.resnets.{j}", b["cin"] if j == 0 else b["cout"], b["cout"])
            if b["spatial"] or b["temporal"]:
                yield from _resampler_shapes(f"encoder.down_blocks.{i}.downsamplers.0", b["cout"], b["spatial"], b["temporal"], up=False)
        yield from _mid_shapes("encoder.mid_block", ech[-1])
        yield "encoder.norm_out.norm.weight", (ech[-1],)
        yield "encoder.norm_out.norm.bias", (ech[-1],)
        yield from _convproj_shapes("encoder.conv_out", ech[-1], c.latent_channels)
        for q in ("quant_conv", "post_quant_conv"):
            yield f"{q}.weight", (c.latent_channels, c.latent_channels, 1, 1, 1)
            yield f"{q}.bias", (c.latent_channels,)
        dch = list(reversed(c.decode_block_out_channels))
        yield from _convproj_shapes("decoder.conv_in", c.latent_channels, dch[0])
        yield from _mid_shapes("decoder.mid_block", dch[0])
        for i, b in enumerate(self.dec_plan):
            for j in range(c.num_layers + 1):
                yield from _resnet_shapes(f"decoder.up_blocks.{i}.resnets.{j}", b["cin"] if j == 0 else b["cout"], b["cout"])
            if b["spatial"] or b["temporal"]:
                yield from _resampler_shapes(f"decoder.up_blocks.{i}.upsamplers.0", b["cout"], b["spatial"], b["temporal"], up=True)
        yield "decoder.norm_out.norm.weight", (dch[-1],)
        yield "decoder.norm_out.norm.bias", (dch[-1],)
        yield from _convproj_shapes("decoder.conv_out", dch[-1], self.inner_dim)

    def _add_param(self, key: str, shape) -> None:
        *path, leaf = key.split(".")
        node = self
        for name in path:
            if name not in node._modules:
                node.add_module(name, _Node())
            node = node._modules[name]
        t = torch.empty(*shape)
        if t.device.type != "meta":
            if len(shape) == 5:
                bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3] * shape[4])
                nn.init.uniform_(t, -bound, bound)
            elif key.endswith("norm.weight"):
                nn.init.ones_(t)
            else:
                nn.init.zeros_(t)
        node.register_parameter(leaf, nn.Parameter(t, requires_grad=False))

    def _p(self, key: str) -> torch.Tensor:
        return self.get_parameter(key)

    def _has(self, key: str) -> bool:
        try:
            self.get_parameter(key)
            return True
        except AttributeError:
            return False

    # ---------------------------------------------------------------------------------------------- packing
    def _pack_key(self):
        w = self._p("quant_conv.weight")
        return (w.device, w.dtype, w.data_ptr(), w._version)

    def _ensure_packed(self) -> Dict[str, torch.Tensor]:
        """Kernel-side copies, built once per (device, dtype): convolution weights [Cout', kt, kh, kw, Cin'] channels-last
        with Cin', Cout' zero-padded to multiples of 64; q|k|v of each attention fused into one [3C, 1,1,1, C] weight; the
        upsampler's (1,3,3) convolution folded into four 2x2 sub-pixel kernels (taps summed in fp32)."""
        key = self._pack_key()
        if self._packed is not None and self._packed_key == key:
            return self._packed
        dev, dt = key[0], key[1]
        if dev.type != "cuda":
            raise RuntimeError("the B200 tokenizer runs on a CUDA device only (no CPU fallback); call .to('cuda') first")
        if dt != BF16:
            raise RuntimeError(f"the B200 tokenizer computes in bfloat16; got parameters in {dt} — call reset_dtype(torch.bfloat16)")
        if not _lib.load().drb_device_supported():
            raise RuntimeError("libdrb200.so targets sm_100a (B200) only; the current CUDA device is not compute capability 10.x")
        P: Dict[str, torch.Tensor] = {}

        def conv_w(w5: torch.Tensor, b: torch.Tensor):
            cout, cin = w5.shape[0], w5.shape[1]
            w = torch.zeros((_pad64(cout), *w5.shape[2:], _pad64(cin)), device=dev, dtype=dt)
            w[:cout, ..., :cin] = w5.permute(0, 2, 3, 4, 1)
            bb = torch.zeros(_pad64(cout), device=dev, dtype=dt)
            bb[:cout] = b
            return w.contiguous(), bb

        for name, p in self.named_parameters():
            if name.endswith(".weight") and p.ndim == 5:
                base = name[:-7]
                P[base + ".w"], P[base + ".b"] = conv_w(p.data, self._p(base + ".bias").data)
        for name in [n[:-len(".to_q.weight")] for n, _ in self.named_parameters() if n.endswith(".to_q.weight")]:
            P[name + ".qkv.w"] = torch.cat([P[f"{name}.{n}.w"] for n in ("to_q", "to_k", "to_v")], dim=0).contiguous()
            P[name + ".qkv.b"] = torch.cat([P[f"{name}.{n}.b"] for n in ("to_q", "to_k", "to_v")], dim=0).contiguous()
        for i, b in enumerate(self.dec_plan):
            if b["spatial"]:
                base = f"decoder.up_blocks.{i}.upsamplers.0.conv2"
                w = P[base + ".w"].float()                      # [Cout, 1, 3, 3, Cin]
                groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}   # parity -> source taps folded into each of the 2 taps
                for py in (0, 1):
                    for px in (0, 1):
                        sub = torch.zeros((w.shape[0], 1, 2, 2, w.shape[4]), device=dev, dtype=torch.float32)
                        for ky, dys in enumerate(groups[py]):
                            for kx, dxs in enumerate(groups[px]):
                                for dy in dys:
                                    for dx in dxs:
                                        sub[:, 0, ky, kx] += w[:, 0, dy, dx]
                        P[f"{base}.sub{py}{px}.w"] = sub.to(dt).contiguous()
        self._packed, self._packed_key = P, key
        return P

    # ---------------------------------------------------------------------------------------------- operators
    @staticmethod
    def _new_stats(T: int, dev) -> torch.Tensor:
        return torch.zeros((T, 2), device=dev, dtype=torch.float64)

    def _conv(self, name: str, x: torch.Tensor, want_stats: bool = False, **kw):
        P = self._packed
        T_out = (kw.get("out_thw") or x.shape[:3])[0]
        stats = self._new_stats(T_out, x.device) if want_stats else None
        y = ops.conv3d_cl(x, P[name + ".w"], P[name + ".b"], stats=stats, **kw)
        return y, stats

    def _conv_projection(self, name: str, x, resid=None, want_stats: bool = True):
        """CosmosConvProjection3d: (1,3,3) spatial then (3,1,1) causal temporal; the skip term rides on the second"""
        h, _ = self._conv(name + ".conv_s", x, pad_h=1, pad_w=1)
        return self._conv(name + ".conv_t", h, want_stats=want_stats, resid=resid,
                          resid_mode=_lib.RES_SAME if resid is not None else _lib.RES_NONE)

    def _norm(self, name: str, x, stats, silu: bool):
        return ops.groupnorm_apply(x, stats, self._p(name + ".norm.weight"), self._p(name + ".norm.bias"), silu)

    def _resnet(self, p: str, x, stats):
        h = self._norm(p + ".norm1", x, stats, True)
        h, hs = self._conv_projection(p + ".conv1", h)
        h = self._norm(p + ".norm2", h, hs, True)
        skip = self._conv(p + ".conv_shortcut", x)[0] if self._has(p + ".conv_shortcut.weight") else x
        return self._conv_projection(p + ".conv2", h, resid=skip)

    def _attention(self, p: str, x, stats, temporal: bool):
        T, H, W, C = x.shape
        h = self._norm(p + ".norm", x, stats, False)
        qkv, _ = self._conv(p + ".qkv", h)                                    # [T, H, W, 3C]
        if temporal:
            o = ops.temporal_attention(qkv)
        elif C == 512 and self.fused_spatial_attention:
            o = ops.spatial_attention_d512(qkv)                               # one flash kernel, no score matrix (the real net)
        else:
            # other widths (reduced test nets): scores GEMM -> row softmax -> P.V GEMM per frame
            n = H * W
            ld = (n + 7) // 8 * 8
            o = torch.empty((T, H, W, C), device=x.device, dtype=BF16)
            flat = qkv.view(T * n, 3 * C)
            if ld != n:   # K rows are read through a [ld, C] window of the frame: give the last frame `ld - n` rows of slack
                flat = torch.zeros((T * n + (ld - n), 3 * C), device=x.device, dtype=BF16)
                flat[:T * n].copy_(qkv.view(T * n, 3 * C))
            s = torch.empty((n, ld), device=x.device, dtype=BF16)
            vt = torch.empty((C, ld), device=x.device, dtype=BF16)
            for f in range(T):
                rows = flat[f * n:(f + 1) * n]
                ops.gemm(rows[:, :C], flat[f * n:f * n + ld, C:2 * C], out=s)
                ops.softmax_rows(s, n, 1.0 / math.sqrt(C))
                ops.transpose(rows[:, 2 * C:], vt)
                ops.gemm(s, vt, out=o[f].view(n, C))
        return self._conv(p + ".to_out.0", o, want_stats=True, resid=x, resid_mode=_lib.RES_SAME)

    def _mid(self, p: str, x, stats):
        x, stats = self._resnet(p + ".resnets.0", x, stats)
        x, stats = self._attention(p + ".attentions.0", x, stats, temporal=False)
        x, stats = self._attention(p + ".temp_attentions.0", x, stats, temporal=True)
        return self._resnet(p + ".resnets.1", x, stats)

    def _downsample(self, p: str, x, spatial: bool, temporal: bool):
        if spatial:
            T, H, W, _ = x.shape
            if H % 2 or W % 2:
                raise ValueError("spatial down-sampling needs even height and width")
            x, _ = self._conv(p + ".conv1", x, stride_hw=2, out_thw=(T, H // 2, W // 2), resid=x, resid_mode=_lib.RES_POOL_HW)
        if temporal:
            T, H, W, _ = x.shape
            if T % 2 == 0:
                raise ValueError("temporal down-sampling needs an odd frame count")
            x, _ = self._conv(p + ".conv2", x, tmode=_lib.TMODE_DOWN2, out_thw=((T - 1) // 2 + 1, H, W), resid=x,
                              resid_mode=_lib.RES_POOL_T)
        return self._conv(p + ".conv3", x, want_stats=True)

    def _upsample(self, p: str, x, spatial: bool, temporal: bool):
        if temporal:
            T, H, W, _ = x.shape
            if T > 1:
                x, _ = self._conv(p + ".conv1", x, tmode=_lib.TMODE_UP2, out_thw=(2 * T - 1, H, W), resid=x,
                                  resid_mode=_lib.RES_FRAME_UP2)
            else:
                x, _ = self._conv(p + ".conv1", x, resid=x, resid_mode=_lib.RES_SAME)
        if spatial:
            T, H, W, C = x.shape
            P = self._packed
            out = torch.empty((T, 2 * H, 2 * W, C), device=x.device, dtype=BF16)
            for py in (0, 1):
                for px in (0, 1):
                    ops.conv3d_cl(x, P[f"{p}.conv2.sub{py}{px}.w"], P[p + ".conv2.b"], pad_h=1 - py, pad_w=1 - px, resid=x,
                                  resid_mode=_lib.RES_NEAREST_UP_HW, out=out, out_scale=2, out_off=(py, px))
            x = out
        return self._conv(p + ".conv3", x, want_stats=True)

    # ---------------------------------------------------------------------------------------------- encode / decode
    def _check(self, x: torch.Tensor, channels: int, what: str) -> None:
        if x.ndim != 5:
            raise ValueError(f"expects a 5D {what} (B, C, T, H, W), but got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("the B200 tokenizer runs on a CUDA device only (no CPU fallback)")
        if x.shape[1] != channels:
            raise ValueError(f"expected {channels} channels, got {x.shape[1]}")

    @torch.no_grad()
    def encode_tensor(self, x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """(B, 3, T, H, W) in [-1, 1] -> latent (B, 16, 1 + (T-1)/8, H/8, W/8), multiplied by `scale` on the way out"""
        c = self.config
        self._check(x, c.in_channels, "input")
        B, _, T, H, W = x.shape
        f = c.spatial_compression_ratio
        if (T - 1) % c.temporal_compression_ratio or H % (2 * f) or W % (2 * f):
            raise ValueError(f"clip {T}x{H}x{W}: frames must be 1 + 8k, height and width multiples of {2 * f}")
        self._ensure_packed()
        outs = []
        for b in range(B):
            h = ops.haar_patch(x[b].to(BF16).contiguous())
            h, st = self._conv_projection("encoder.conv_in", h)
            for i, blk in enumerate(self.enc_plan):
                for j in range(c.num_layers):
                    h, st = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", h, st)
                if blk["spatial"] or blk["temporal"]:
                    h, st = self._downsample(f"encoder.down_blocks.{i}.downsamplers.0", h, blk["spatial"], blk["temporal"])
            h, st = self._mid("encoder.mid_block", h, st)
            h = self._norm("encoder.norm_out", h, st, True)
            h, _ = self._conv_projection("encoder.conv_out", h, want_stats=False)
            h, _ = self._conv("quant_conv", h)
            outs.append(ops.cl_to_planar(h, c.latent_channels, scale))
        return torch.stack(outs, dim=0)

    @torch.no_grad()
    def decode_tensor(self, z: torch.Tensor, scale: float = 1.0, to_u8: bool = False, normalize_normal: bool = False) -> torch.Tensor:
        """latent (B, 16, t, h, w), multiplied by `scale` on the way in -> (B, 3, 1 + 8(t-1), 8h, 8w); with `to_u8` the
        last stage stores the post-processed uint8 frames (B, T, H, W, 3) instead (diffusion_renderer_pipeline.py:300-318
        fused into the inverse-Haar store)"""
        c = self.config
        self._check(z, c.latent_channels, "latent")
        self._ensure_packed()
        outs = []
        for b in range(z.shape[0]):
            h = ops.planar_to_cl(z[b].to(BF16).contiguous(), _pad64(c.latent_channels), scale)
            h, _ = self._conv("post_quant_conv", h)
            h, st = self._conv_projection("decoder.conv_in", h)
            h, st = self._mid("decoder.mid_block", h, st)
            for i, blk in enumerate(self.dec_plan):
                for j in range(c.num_layers + 1):
                    h, st = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", h, st)
                if blk["spatial"] or blk["temporal"]:
                    h, st = self._upsample(f"decoder.up_blocks.{i}.upsamplers.0", h, blk["spatial"], blk["temporal"])
            h = self._norm("decoder.norm_out", h, st, True)
            h, _ = self._conv_projection("decoder.conv_out", h, want_stats=False)
            outs.append(ops.haar_unpatch_u8(h, normalize_normal) if to_u8 else ops.haar_unpatch(h))
        return torch.stack(outs, dim=0)

    # the diffusers call surface the reference wrapper uses (CleanVAE.py:50-51, 59-60)
    def encode(self, x: torch.Tensor):
        z = self.encode_tensor(x)
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda: z, mode=lambda: z))

    def decode(self, z: torch.Tensor):
        return SimpleNamespace(sample=self.decode_tensor(z))

    @classmethod
    def from_pretrained(cls, model_path: str) -> "AutoencoderKLCosmos":
        """`config.json` + `diffusion_pytorch_model.safetensors` of a diffusers checkpoint directory."""
        cfg_path = os.path.join(model_path, "config.json")
        config = {}
        if os.path.isfile(cfg_path):
            with open(cfg_path) as f:
                config = json.load(f)
        model = cls(**config)
        for fname in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin"):
            path = os.path.join(model_path, fname)
            if os.path.isfile(path):
                if fname.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(path)
                else:
                    sd = torch.load(path, map_location="cpu", weights_only=True)
                model.load_state_dict(sd, strict=True)
                return model
        raise FileNotFoundError(f"no diffusion_pytorch_model.safetensors / .bin under {model_path}")


class CleanVAE:
    """The reference wrapper (CleanVAE.py:9-67), on top of this module's AutoencoderKLCosmos."""

    def __init__(self, model_path: Optional[str] = None, model: Optional[AutoencoderKLCosmos] = None):
        if model is None:
            if model_path is None:
                raise ValueError("CleanVAE needs a model_path (or a ready AutoencoderKLCosmos via model=)")
            model = AutoencoderKLCosmos.from_pretrained(model_path)
        if model is None:
            raise ValueError(f"Failed to load VAE model from {model_path}")
        self.model = model
        self.config = self.model.config
        self.spatial_compression_factor = self.model.config.spatial_compression_ratio
        self.latent_ch = self.config.latent_channels
        self.temporal_compression_factor = 8

    def get_latent_num_frames(self, num_pixel_frames: int) -> int:
        if num_pixel_frames == 1:
            return 1
        return (num_pixel_frames - 1) // self.temporal_compression_factor + 1

    def get_pixel_num_frames(self, num_latent_frames: int) -> int:
        if num_latent_frames == 1:
            return 1
        return (num_latent_frames - 1) * self.temporal_compression_factor + 1

    @torch.no_grad()
    def encode(self, state_5d: torch.Tensor) -> torch.Tensor:
        if state_5d.ndim != 5:
            raise ValueError(f"CleanVAE expects a 5D input (B, C, T, H, W), but got {state_5d.shape}")
        return self.model.encode(state_5d).latent_dist.sample()

    @torch.no_grad()
    def decode(self, latent_5d: torch.Tensor) -> torch.Tensor:
        if latent_5d.ndim != 5:
            raise ValueError(f"CleanVAE expects a 5D latent (B, C, T, H, W), but got {latent_5d.shape}")
        return self.model.decode(latent_5d).sample

    # fused-scale forms used by CleanDiffusionRendererModel.encode / .decode (model_diffusion_renderer.py:138-156): the
    # sigma_data factor rides on the boundary layout kernels instead of a separate tensor op
    @torch.no_grad()
    def encode_scaled(self, state_5d: torch.Tensor, scale: float) -> torch.Tensor:
        return self.model.encode_tensor(state_5d, scale)

    @torch.no_grad()
    def decode_scaled(self, latent_5d: torch.Tensor, scale: float) -> torch.Tensor:
        return self.model.decode_tensor(latent_5d, scale)

    @torch.no_grad()
    def decode_u8(self, latent_5d: torch.Tensor, scale: float = 1.0, normalize_normal: bool = False) -> torch.Tensor:
        """decode + the pipeline's uint8 post-process in one go: latent (B,16,t,h,w) -> uint8 (B,T,H,W,3) on the device"""
        if latent_5d.ndim != 5:
            raise ValueError(f"CleanVAE expects a 5D latent (B, C, T, H, W), but got {latent_5d.shape}")
        return self.model.decode_tensor(latent_5d, scale, to_u8=True, normalize_normal=normalize_normal)

    # Chunked form for clips longer than one tokenizer window — the behaviour of the upstream chunking tokenizer that the
    # reference carries as (unused) `BasePretrainedVideoTokenizer.encode / .decode` (pretrained_vae.py:389-440): the clip is
    # cut into windows of `pixel_chunk_duration` frames (121 for CV8x8x8: 16 latent frames each), every window is encoded
    # / decoded on its own (so the first frame of every window is again the tokenizer's "image" frame) and the results
    # are concatenated in time.
    @staticmethod
    def _latent_chunk(pixel_chunk_duration: int, factor: int = 8) -> int:
        if (pixel_chunk_duration - 1) % factor:
            raise ValueError(f"pixel_chunk_duration {pixel_chunk_duration} must be 1 + a multiple of {factor}")
        return (pixel_chunk_duration - 1) // factor + 1

    def _chunk_stats(self, frames: int, repeat: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """bf16 (mean, std), one per (chunk, channel, latent frame) row: the config's latents_mean / latents_std viewed as
        (C, -1) and cut to the chunk's frame count (the layout diffusers' Cosmos pipelines use), repeated per chunk"""
        mean, std = getattr(self.config, "latents_mean", None), getattr(self.config, "latents_std", None)
        if mean is None or std is None:
            raise ValueError("the tokenizer config carries no latents_mean / latents_std (VAE_config.json:21-536)")
        C = self.latent_ch
        m = torch.tensor(mean, dtype=torch.float32).reshape(C, -1)
        sd = torch.tensor(std, dtype=torch.float32).reshape(C, -1)
        if m.shape[1] < frames:
            raise ValueError(f"the config holds statistics for {m.shape[1]} latent frames per chunk, the chunk has {frames}")
        to = lambda t: t[:, :frames].reshape(1, -1).repeat(repeat, 1).reshape(-1).to(device=device, dtype=torch.bfloat16).contiguous()
        return to(m), to(sd)

    @torch.no_grad()
    def encode_chunked(self, state_5d: torch.Tensor, pixel_chunk_duration: int = 121, normalize: bool = False,
                       max_enc_batch_size: int = 8) -> torch.Tensor:
        """`normalize`: apply the per-chunk latent statistics, (z - mean) / std per (channel, latent frame)
        (pretrained_vae.py:142); at most `max_enc_batch_size` chunks go through the tokenizer per call (:389-405)."""
        if state_5d.ndim != 5:
            raise ValueError(f"CleanVAE expects a 5D input (B, C, T, H, W), but got {state_5d.shape}")
        B, C, T, H, W = state_5d.shape
        self._latent_chunk(pixel_chunk_duration, self.temporal_compression_factor)
        if T % pixel_chunk_duration:
            raise ValueError(f"Temporal dimension {T} is not divisible by chunk_length {pixel_chunk_duration}")
        n = T // pixel_chunk_duration
        chunks = state_5d.reshape(B, C, n, pixel_chunk_duration, H, W).permute(0, 2, 1, 3, 4, 5).reshape(B * n, C, pixel_chunk_duration, H, W)
        step = max(1, int(max_enc_batch_size))
        z = torch.cat([self.encode(chunks[i:i + step]) for i in range(0, B * n, step)], dim=0)      # (B*n, 16, t, h, w)
        _, c, t, h, w = z.shape
        if normalize:
            mean, std = self._chunk_stats(t, B * n, z.device)
            z = ops.latent_normalize(z.contiguous(), mean, std, decode=False)
        return z.reshape(B, n, c, t, h, w).permute(0, 2, 1, 3, 4, 5).reshape(B, c, n * t, h, w)

    @torch.no_grad()
    def decode_chunked(self, latent_5d: torch.Tensor, pixel_chunk_duration: int = 121, normalize: bool = False,
                       max_dec_batch_size: int = 4) -> torch.Tensor:
        """`normalize`: undo the per-chunk latent statistics first, z * std + mean (pretrained_vae.py:150); at most
        `max_dec_batch_size` chunks go through the tokenizer per call (:424-433)."""
        if latent_5d.ndim != 5:
            raise ValueError(f"CleanVAE expects a 5D latent (B, C, T, H, W), but got {latent_5d.shape}")
        B, c, T, h, w = latent_5d.shape
        lc = self._latent_chunk(pixel_chunk_duration, self.temporal_compression_factor)
        if T % lc:
            raise ValueError(f"Temporal dimension {T} is not divisible by chunk_length {lc}")
        n = T // lc
        chunks = latent_5d.reshape(B, c, n, lc, h, w).permute(0, 2, 1, 3, 4, 5).reshape(B * n, c, lc, h, w)
        if normalize:
            mean, std = self._chunk_stats(lc, B * n, chunks.device)
            chunks = ops.latent_normalize(chunks.to(torch.bfloat16).contiguous(), mean, std, decode=True)
        step = max(1, int(max_dec_batch_size))
        y = torch.cat([self.decode(chunks[i:i + step]) for i in range(0, B * n, step)], dim=0)      # (B*n, 3, pixel_chunk_duration, H, W)
        _, C, t, H, W = y.shape
        return y.reshape(B, n, C, t, H, W).permute(0, 2, 1, 3, 4, 5).reshape(B, C, n * t, H, W)

    def to(self, device):
        self.model = self.model.to(device)
        return self

    def reset_dtype(self, dtype: torch.dtype):
        self.model.to(dtype)
