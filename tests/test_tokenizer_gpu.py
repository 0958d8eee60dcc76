"""GPU parity of the B200 tokenizer (CleanVAE.py + conv.cu / tokenizer.cu) against oracle/vae_oracle.py on the same GPU.

The oracle itself is PARITY UNPINNED (diffusers >= 0.34 is not installable here; see oracle/vae_oracle.py), so these
tests pin the kernels to the restated algorithm, not to upstream bits.  Tolerances: single operators — relative L2
<= 4e-3 against an fp32 evaluation of the same bf16 inputs (bf16 output rounding alone is ~2e-3); whole encode / decode —
error against the fp32 oracle no worse than 2x the bf16 oracle's own error + 2e-3, and decoded frames >= 40 dB PSNR
(BASELINE.json north_star)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as vo
from tests.util import psnr_u8, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def gen(seed=0):
    return torch.Generator(device=DEV).manual_seed(seed)


def cl(x):     # [1,C,T,H,W] -> [T,H,W,C] contiguous
    return x[0].permute(1, 2, 3, 0).contiguous()


def ncthw(x):  # [T,H,W,C] -> [1,C,T,H,W]
    return x.permute(3, 0, 1, 2).unsqueeze(0)


def rand_conv(cout, cin, kt, kh, kw, seed):
    g = gen(seed)
    w = ((torch.rand(cout, cin, kt, kh, kw, device=DEV, generator=g) * 2 - 1) / math.sqrt(cin * kt * kh * kw)).bfloat16()
    b = (0.1 * torch.randn(cout, device=DEV, generator=g)).bfloat16()
    return w, b


def wcl(w):    # torch [Cout,Cin,kt,kh,kw] -> kernel [Cout,kt,kh,kw,Cin]
    return w.permute(0, 2, 3, 4, 1).contiguous()


def sd32(w, b, name="c"):
    return {f"{name}.weight": w.float(), f"{name}.bias": b.float()}


def stats_of(y):
    y = y.float()
    return torch.stack([y.sum(dim=(1, 2, 3)), (y * y).sum(dim=(1, 2, 3))], dim=1)


@pytest.mark.parametrize("shape,k,pad", [
    ((3, 8, 16, 64, 256), (1, 3, 3), 1),       # one exact tile per frame, N = 256
    ((5, 20, 28, 128, 64), (1, 3, 3), 1),      # ragged tiles, N = 64
    ((4, 12, 40, 192, 128), (3, 1, 1), 0),     # causal temporal taps, 3 channel chunks
    ((2, 9, 10, 64, 512), (1, 1, 1), 0),       # 1x1x1, two n-tiles
    ((1, 8, 16, 64, 1536), (1, 1, 1), 0),      # fused q|k|v projection width
    ((6, 4, 4, 64, 48), (3, 1, 1), 0),         # frame smaller than the tile, N = 48
])
def test_conv3d_matches_torch(shape, k, pad):
    from drb200 import ops
    T, H, W, cin, cout = shape
    w, b = rand_conv(cout, cin, *k, seed=1)
    x = torch.randn(1, cin, T, H, W, device=DEV, generator=gen(2)).bfloat16()
    stats = torch.zeros(T, 2, device=DEV, dtype=torch.float64)
    got = ops.conv3d_cl(cl(x), wcl(w), b, pad_h=pad, pad_w=pad, stats=stats)
    ref = cl(vo.causal_conv3d(sd32(w, b), "c", x.float(), padding=pad))
    assert got.shape == ref.shape
    assert rel_l2(got, ref) <= 4e-3
    assert torch.allclose(stats, stats_of(got).double(), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("T,H,W,c,k", [
    (3, 16, 16, 128, (3, 1, 1)),     # whole tiles, two quarter buffers per tile
    (2, 9, 21, 96, (3, 1, 1)),       # ragged tiles, a half-filled quarter buffer
    (2, 10, 12, 384, (1, 1, 1)),     # second n-tile only half inside Cout
    (3, 8, 16, 192, (1, 3, 3)),      # three quarter buffers per tile: the buffer ring wraps inside a tile
    (5, 24, 32, 256, (3, 1, 1)),     # more tiles than CTAs hold at once: skip vectors requested across the tile boundary
    (2, 16, 16, 512, (1, 3, 3)),     # 72 k-blocks: the single-stage epilogue
])
def test_conv3d_skip_term_and_stats(T, H, W, c, k):
    """skip term + GroupNorm sums through both epilogue flavours (the split one up to 36 k-blocks per tile)"""
    from drb200 import ops
    cin = c if c % 64 == 0 else 64               # Cin comes in 64-channel chunks; Cout in multiples of 16
    w, b = rand_conv(c, cin, *k, seed=3)
    pad = 1 if k[1] == 3 else 0
    x = torch.randn(1, cin, T, H, W, device=DEV, generator=gen(4)).bfloat16()
    skip = torch.randn(1, c, T, H, W, device=DEV, generator=gen(5)).bfloat16()
    stats = torch.zeros(T, 2, device=DEV, dtype=torch.float64)
    # the output sits between two guard regions: a store outside [T, H, W, c] (ragged tiles, half-filled buffers) would show
    n_out, guard = T * H * W * c, 4096
    buf = torch.full((n_out + 2 * guard,), 777.0, device=DEV, dtype=torch.bfloat16)
    got = ops.conv3d_cl(cl(x), wcl(w), b, pad_h=pad, pad_w=pad, resid=cl(skip), resid_mode=1, stats=stats,
                        out=buf[guard:guard + n_out].view(T, H, W, c))
    assert (buf[:guard] == 777.0).all() and (buf[guard + n_out:] == 777.0).all()
    conv = cl(vo.causal_conv3d(sd32(w, b), "c", x.float(), padding=pad))
    ref = conv.bfloat16().float() + cl(skip).float()      # the reference rounds the convolution output before the add
    assert rel_l2(got, ref) <= 4e-3
    assert torch.allclose(stats, stats_of(got).double(), rtol=1e-3, atol=1e-2)
    # the same convolution without the skip term, then added in torch: bit-identical (one rounding each, same order)
    plain = ops.conv3d_cl(cl(x), wcl(w), b, pad_h=pad, pad_w=pad)
    assert torch.equal(got, (plain.float() + cl(skip).float()).bfloat16())


@pytest.mark.parametrize("T,H,W", [(7, 16, 32), (1, 8, 8), (5, 12, 20)])
def test_downsample_matches_oracle(T, H, W):
    """CosmosDownsample3d = strided (1,3,3) conv + 2x2 avg pool, strided causal (3,1,1) conv + 2-frame avg pool, 1x1x1."""
    from drb200 import _lib, ops
    c = 64
    ws = [rand_conv(c, c, 1, 3, 3, 6), rand_conv(c, c, 3, 1, 1, 7), rand_conv(c, c, 1, 1, 1, 8)]
    sd = {}
    for i, (w, b) in enumerate(ws):
        sd.update(sd32(w, b, f"d.conv{i + 1}"))
    x = torch.randn(1, c, T, H, W, device=DEV, generator=gen(9)).bfloat16()
    h = cl(x)
    h = ops.conv3d_cl(h, wcl(ws[0][0]), ws[0][1], stride_hw=2, out_thw=(T, H // 2, W // 2), resid=h, resid_mode=_lib.RES_POOL_HW)
    mid_ref = vo.downsample({**sd, "d.conv3.weight": torch.eye(c, device=DEV).reshape(c, c, 1, 1, 1), "d.conv3.bias": torch.zeros(c, device=DEV)},
                            "d", x.float(), True, False)
    assert rel_l2(h, cl(mid_ref)) <= 4e-3
    h = ops.conv3d_cl(h, wcl(ws[1][0]), ws[1][1], tmode=_lib.TMODE_DOWN2, out_thw=((T - 1) // 2 + 1, H // 2, W // 2), resid=h,
                      resid_mode=_lib.RES_POOL_T)
    h = ops.conv3d_cl(h, wcl(ws[2][0]), ws[2][1])
    ref = cl(vo.downsample(sd, "d", x.float(), True, True))
    assert h.shape == ref.shape
    assert rel_l2(h, ref) <= 6e-3


@pytest.mark.parametrize("T,H,W", [(4, 8, 16), (1, 8, 8), (3, 6, 10)])
def test_upsample_matches_oracle(T, H, W):
    """CosmosUpsample3d through the product's own sequencing (temporal conv on the interleaved frames, four sub-pixel
    2x2 convolutions for nearest-x2 + (1,3,3), 1x1x1)."""
    from drb200.CleanVAE import AutoencoderKLCosmos
    model = AutoencoderKLCosmos(encoder_block_out_channels=(64, 64, 64, 64), decode_block_out_channels=(64, 64, 64, 64))
    p = "decoder.up_blocks.1.upsamplers.0"
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if n.startswith(p) and n.endswith("bias"):
                prm.copy_(0.1 * torch.randn(prm.shape, generator=g))
    model = model.to(DEV).to(torch.bfloat16)
    model._ensure_packed()
    sd = {k: v.float() for k, v in model.state_dict().items()}
    x = torch.randn(1, 64, T, H, W, device=DEV, generator=gen(12)).bfloat16()
    got, stats = model._upsample(p, cl(x), True, True)
    ref = cl(vo.upsample(sd, p, x.float(), True, True))
    assert got.shape == ref.shape
    assert rel_l2(got, ref) <= 6e-3
    assert torch.allclose(stats, stats_of(got).double(), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("T,H,W", [(9, 32, 64), (1, 16, 16), (5, 8, 260)])
def test_haar_patch_and_unpatch(T, H, W):
    from drb200 import ops
    x = (torch.rand(1, 3, T, H, W, device=DEV, generator=gen(13)) * 2 - 1).bfloat16()
    got = ops.haar_patch(x[0].contiguous())
    ref = cl(vo.haar_patch(x.float()))
    assert got.shape == ref.shape
    assert (got.float() - ref).abs().max() <= 2 ** -8 * ref.abs().max()          # one rounding of an exact fp32 result
    back = ops.haar_unpatch(got)
    assert back.shape == (3, T, H, W)
    assert (back.float() - x[0].float()).abs().max() <= 2 ** -6
    coef = torch.randn(got.shape, device=DEV, generator=gen(14)).bfloat16()
    ref_un = vo.haar_unpatch(ncthw(coef.float()))[0]
    got_un = ops.haar_unpatch(coef)
    assert rel_l2(got_un, ref_un) <= 3e-3


@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm_apply_and_frame_stats(silu):
    from drb200 import ops
    T, H, W, c = 3, 10, 12, 128
    x = (torch.randn(1, c, T, H, W, device=DEV, generator=gen(15)) * 2 + 0.5).bfloat16()
    gamma = (1 + 0.1 * torch.randn(c, device=DEV, generator=gen(16))).bfloat16()
    beta = (0.1 * torch.randn(c, device=DEV, generator=gen(17))).bfloat16()
    xc = cl(x)
    stats = ops.frame_stats(xc)
    assert torch.allclose(stats, stats_of(xc).double(), rtol=1e-4)
    got = ops.groupnorm_apply(xc, stats, gamma, beta, silu)
    ref = vo.causal_group_norm({"n.norm.weight": gamma, "n.norm.bias": beta}, "n", x)
    if silu:
        ref = F.silu(ref)
    ref = cl(ref)
    bad = (got.float() - ref.float()).abs() > 2 ** -6 * ref.float().abs().clamp_min(0.25)
    assert not bad.any()
    assert (got == ref).float().mean() > 0.97      # incl. torch's bf16-rounded mean / rstd (see tokenizer.cu)


def test_groupnorm_apply_propagates_nan_and_inf():
    """the integer round-to-nearest-even of the normalised value must not carry a NaN (0x7fffffff on the GPU) into -0"""
    from drb200 import ops
    T, H, W, c = 1, 8, 8, 64
    x = torch.randn(T, H, W, c, device=DEV, generator=gen(31)).bfloat16()
    stats = ops.frame_stats(x)
    x[0, 1, 2, 3] = float("nan")
    x[0, 4, 5, 6] = float("inf")
    ones, zeros = torch.ones(c, device=DEV, dtype=torch.bfloat16), torch.zeros(c, device=DEV, dtype=torch.bfloat16)
    for silu in (False, True):
        got = ops.groupnorm_apply(x, stats, ones, zeros, silu)
        assert torch.isnan(got[0, 1, 2, 3]) and not torch.isfinite(got[0, 4, 5, 6])    # SiLU(+Inf) may come out as NaN
        assert torch.isfinite(got.float()).sum() == got.numel() - 2


def test_softmax_transpose_temporal_attention():
    from drb200 import ops
    n, cols, ld = 70, 77, 80
    s = (torch.randn(n, ld, device=DEV, generator=gen(18)) * 4).bfloat16()
    ref = torch.softmax(s[:, :cols].float() * 0.3, dim=-1)
    ops.softmax_rows(s, cols, 0.3)
    assert (s[:, cols:] == 0).all()
    assert (s[:, :cols].float() - ref).abs().max() <= 2 ** -8
    wide = (torch.randn(3, 17008, device=DEV, generator=gen(23)) * 3).bfloat16()          # > 16384 columns: the three-pass path
    ref_w = torch.softmax(wide[:, :17001].float() * 0.2, dim=-1)
    ops.softmax_rows(wide, 17001, 0.2)
    assert (wide[:, 17001:] == 0).all() and (wide[:, :17001].float() - ref_w).abs().max() <= 2 ** -8 * ref_w.max()
    a = torch.randn(45, 3 * 64, device=DEV, generator=gen(19)).bfloat16()
    out = torch.full((64, 48), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.transpose(a[:, 64:128], out)
    assert torch.equal(out[:, :45], a[:, 64:128].t()) and (out[:, 45:] == 0).all()
    T, H, W, c = 6, 5, 7, 128
    qkv = torch.randn(T, H, W, 3 * c, device=DEV, generator=gen(20)).bfloat16()
    got = ops.temporal_attention(qkv)
    q, k, v = (t.permute(1, 2, 0, 3).reshape(H * W, 1, T, c).float() for t in qkv.split(c, dim=-1))
    mask = torch.tril(torch.ones(T, T, device=DEV)).bool()
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=mask).reshape(H, W, T, c).permute(2, 0, 1, 3)
    assert rel_l2(got, ref) <= 4e-3


def _product_vae(dims, seed):
    from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE
    sd = vo.make_vae_state_dict(dims, seed=seed)
    model = AutoencoderKLCosmos(encoder_block_out_channels=dims.encoder_block_out_channels,
                                decode_block_out_channels=dims.decode_block_out_channels)
    model.load_state_dict(sd, strict=True)
    vae = CleanVAE(model=model)
    vae.to(DEV)
    vae.reset_dtype(torch.bfloat16)
    return vae, {k: v.to(DEV) for k, v in sd.items()}


@pytest.mark.parametrize("dims,thw", [
    (vo.SMALL_VAE, (9, 64, 96)),       # every code path at 1/8 width; ragged spatial tiles; H*W/64 = 96 tokens
    (vo.SMALL_VAE, (1, 32, 32)),       # image path (T = 1)
    (vo.SMALL_VAE, (17, 48, 80)),      # 3 latent frames, 60 spatial tokens (ld padding of the score matrix)
    (vo.FULL_VAE, (9, 64, 64)),        # the real channel widths
])
def test_encode_decode_match_oracle(dims, thw):
    vae, sd = _product_vae(dims, seed=7)
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    T, H, W = thw
    x = (torch.rand(1, 3, T, H, W, device=DEV, generator=gen(21)) * 2 - 1).bfloat16()
    with torch.no_grad():
        z = vae.encode(x)
        z32 = vo.encode({k: v.bfloat16().float() for k, v in sd.items()}, dims, x.float())
        z16 = vo.encode(sd16, dims, x)
    assert z.shape == z32.shape == (1, 16, (T - 1) // 8 + 1, H // 8, W // 8) and z.dtype == torch.bfloat16
    floor, err = rel_l2(z16, z32), rel_l2(z, z32)
    print(f"\nencode {thw}: product-vs-fp32 {err:.3e}   bf16-oracle-vs-fp32 {floor:.3e}")
    assert err <= 2 * floor + 2e-3
    with torch.no_grad():
        zin = z32.bfloat16()
        y = vae.decode(zin)
        y32 = vo.decode({k: v.bfloat16().float() for k, v in sd.items()}, dims, zin.float())
        y16 = vo.decode(sd16, dims, zin)
    assert y.shape == y32.shape == (1, 3, T, H, W)
    floor, err = rel_l2(y16, y32), rel_l2(y, y32)
    print(f"decode {thw}: product-vs-fp32 {err:.3e}   bf16-oracle-vs-fp32 {floor:.3e}")
    assert err <= 2 * floor + 2e-3

    def u8(v):   # diffusion_renderer_pipeline.py:300-318 in plain torch
        return (((1 + v.float()).clamp(0, 2) / 2) * 255).to(torch.uint8).cpu().numpy()
    scale = 1.0 / max(1.0, float(y32.abs().max()))      # random-init decoders are not confined to [-1, 1]
    p = psnr_u8(u8(y * scale), u8(y32 * scale))
    print(f"decoded frames PSNR vs fp32 oracle: {p:.1f} dB")
    assert p >= 40.0


def test_model_encode_decode_use_the_b200_tokenizer():
    """CleanDiffusionRendererModel.encode / .decode (model_diffusion_renderer.py:138-156) with the sigma_data factor"""
    from oracle.weights import MICRO_INVERSE
    from tests.util import build_product_model
    vae, sd = _product_vae(vo.SMALL_VAE, seed=8)
    model, _ = build_product_model(MICRO_INVERSE, "inverse", seed=3, vae=vae)
    x = (torch.rand(1, 3, 9, 32, 48, device=DEV, generator=gen(22)) * 2 - 1).bfloat16()
    with torch.no_grad():
        z = model.encode(x)
        assert torch.equal(z, vae.encode(x) * 0.5)
        y = model.decode(z)
        assert torch.equal(y, vae.decode(z / 0.5))
    with pytest.raises(ValueError):
        vae.encode(x[0])
    with pytest.raises(ValueError):
        vae.decode(z[0])


def test_chunked_encode_decode_equals_per_window_calls():
    """encode_chunked / decode_chunked (the upstream chunking tokenizer's semantics, pretrained_vae.py:389-440): windows of
    `pixel_chunk_duration` frames are tokenised independently and concatenated in time"""
    vae, sd = _product_vae(vo.SMALL_VAE, seed=7)
    pcd = 17                                             # 3 latent frames per window
    x = (torch.rand(1, 3, 2 * pcd, 32, 48, device=DEV, generator=gen(31)) * 2 - 1).bfloat16()
    z = vae.encode_chunked(x, pixel_chunk_duration=pcd)
    assert z.shape == (1, 16, 6, 4, 6)
    assert torch.equal(z[:, :, :3], vae.encode(x[:, :, :pcd])) and torch.equal(z[:, :, 3:], vae.encode(x[:, :, pcd:]))
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    ref = torch.cat([vo.encode(sd16, vo.SMALL_VAE, x[:, :, i * pcd:(i + 1) * pcd]) for i in range(2)], dim=2)
    assert rel_l2(z, ref) <= 3e-2
    y = vae.decode_chunked(z, pixel_chunk_duration=pcd)
    assert y.shape == (1, 3, 2 * pcd, 32, 48)
    assert torch.equal(y[:, :, pcd:], vae.decode(z[:, :, 3:]))
    with pytest.raises(ValueError):
        vae.encode_chunked(x[:, :, :30], pixel_chunk_duration=pcd)
    with pytest.raises(ValueError):
        vae.decode_chunked(z[:, :, :4], pixel_chunk_duration=pcd)
    with pytest.raises(ValueError):
        vae.encode_chunked(x, pixel_chunk_duration=16)


def test_chunked_per_chunk_latent_normalisation():
    """opt-in per-chunk mean / std of the upstream chunking tokenizer (pretrained_vae.py:142,150) with the statistics layout
    of VAE_config.json:21-536 (16 channels x 16 latent frames): bit-equal to the reference's bf16 tensor arithmetic, and
    decode_chunked(normalize=True) undoes encode_chunked(normalize=True)"""
    import json
    import os
    from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE
    cfg_path = os.path.join(os.path.dirname(__file__), "golden", "vae_latent_stats.json")
    with open(cfg_path) as f:
        stats = json.load(f)                              # latents_mean / latents_std copied from the reference config
    assert len(stats["latents_mean"]) == len(stats["latents_std"]) == 256
    sd = vo.make_vae_state_dict(vo.SMALL_VAE, seed=7)
    model = AutoencoderKLCosmos(encoder_block_out_channels=vo.SMALL_VAE.encoder_block_out_channels,
                                decode_block_out_channels=vo.SMALL_VAE.decode_block_out_channels, **stats)
    model.load_state_dict(sd, strict=True)
    vae = CleanVAE(model=model)
    vae.to(DEV)
    vae.reset_dtype(torch.bfloat16)
    pcd = 17                                              # 3 latent frames per chunk
    x = (torch.rand(1, 3, 2 * pcd, 32, 48, device=DEV, generator=gen(33)) * 2 - 1).bfloat16()
    z_raw = vae.encode_chunked(x, pixel_chunk_duration=pcd)
    z = vae.encode_chunked(x, pixel_chunk_duration=pcd, normalize=True, max_enc_batch_size=1)
    mean = torch.tensor(stats["latents_mean"], device=DEV).view(1, 16, -1, 1, 1)[:, :, :3].bfloat16()
    std = torch.tensor(stats["latents_std"], device=DEV).view(1, 16, -1, 1, 1)[:, :, :3].bfloat16()
    want = torch.cat([(z_raw[:, :, i * 3:(i + 1) * 3] - mean) / std for i in range(2)], dim=2)     # bf16 tensor ops, as :142
    assert torch.equal(z, want)
    back = torch.cat([z[:, :, i * 3:(i + 1) * 3] * std + mean for i in range(2)], dim=2)           # :150
    y = vae.decode_chunked(z, pixel_chunk_duration=pcd, normalize=True, max_dec_batch_size=1)
    assert torch.equal(y, vae.decode_chunked(back, pixel_chunk_duration=pcd))
    plain, _ = _product_vae(vo.SMALL_VAE, seed=7)         # a config without statistics refuses
    with pytest.raises(ValueError):
        plain.encode_chunked(x, pixel_chunk_duration=pcd, normalize=True)


@pytest.mark.parametrize("T,H,W,gain", [(2, 8, 8, 1.0), (3, 12, 20, 1.0), (1, 44, 80, 1.0), (2, 24, 40, 6.0)])
def test_fused_spatial_attention_d512(T, H, W, gain):
    """drb_spatial_attention_d512 (one head of dim 512 over the H*W tokens of a frame, flash form) against fp32 SDPA and
    against the GEMM -> softmax -> GEMM form it replaces; `gain` 6 gives logits of +-40 whose row maxima keep moving"""
    import torch.nn.functional as F
    from drb200 import ops
    g = gen(41)
    qkv = torch.randn(T, H, W, 1536, device=DEV, generator=g)
    qkv[..., :1024] *= gain ** 0.5 * torch.linspace(0.5, 1.5, H * W, device=DEV).reshape(1, H, W, 1)
    qkv = qkv.bfloat16()
    out = ops.spatial_attention_d512(qkv)
    n = H * W
    flat = qkv.float().reshape(T, n, 1536)
    ref = F.scaled_dot_product_attention(flat[:, None, :, :512], flat[:, None, :, 512:1024], flat[:, None, :, 1024:])[:, 0]
    assert out.shape == (T, H, W, 512) and torch.isfinite(out.float()).all()
    err = rel_l2(out.reshape(T, n, 512), ref)
    print(f"\nspatial attention {T}x{H}x{W} gain {gain}: rel-L2 vs fp32 SDPA {err:.3e}")
    assert err <= 5e-3
    # the unfused form (scores GEMM -> row softmax -> P.V GEMM) rounds the scores to bf16 first: close, not identical
    ld = (n + 7) // 8 * 8
    s = torch.empty((n, ld), device=DEV, dtype=torch.bfloat16)
    vt = torch.empty((512, ld), device=DEV, dtype=torch.bfloat16)
    rows = qkv.reshape(T * n, 1536)
    if n % 8 == 0:
        ops.gemm(rows[:n, :512], rows[:n, 512:1024], out=s)
        ops.softmax_rows(s, n, 1.0 / 512 ** 0.5)
        ops.transpose(rows[:n, 1024:], vt)
        old = ops.gemm(s, vt)
        assert rel_l2(out.reshape(T, n, 512)[0], old) <= 2e-2 * max(1.0, gain)


def test_full_width_tokenizer_with_and_without_the_fused_spatial_attention():
    """the whole encode with the mid-block attention as one flash kernel vs the GEMM -> softmax -> GEMM form: same result up
    to the bf16 rounding of the scores the unfused form adds"""
    vae, sd = _product_vae(vo.FULL_VAE, seed=9)
    x = (torch.rand(1, 3, 9, 128, 192, device=DEV, generator=gen(44)) * 2 - 1).bfloat16()
    z = vae.encode(x)
    vae.model.fused_spatial_attention = False
    z_unfused = vae.encode(x)
    vae.model.fused_spatial_attention = True
    assert torch.isfinite(z.float()).all() and rel_l2(z, z_unfused) <= 1e-2
    z32 = vo.encode({k: v.bfloat16().float() for k, v in sd.items()}, vo.FULL_VAE, x.float())
    assert rel_l2(z, z32) <= rel_l2(z_unfused, z32) + 2e-3


@pytest.mark.parametrize("normalize", [False, True])
def test_fused_uint8_store_equals_decode_then_postprocess(normalize):
    """drb_haar_unpatch_u8 (the post-process of diffusion_renderer_pipeline.py:300-318 fused into the
    tokenizer's last store) is bit-identical to inverse Haar -> planar bf16 video -> drb_postprocess_u8"""
    from drb200 import ops
    g = gen(51)
    for Tp, Hp, Wp in ((1, 3, 5), (3, 6, 70), (2, 4, 64)):          # ragged blocks of 64 pixels, image path (Tp = 1)
        x = (torch.randn(Tp, Hp, Wp, 192, device=DEV, generator=g) * 0.35).bfloat16()
        want = ops.postprocess_u8(ops.haar_unpatch(x), normalize)
        got = ops.haar_unpatch_u8(x, normalize)
        assert got.shape == want.shape == (4 * Tp - 3, 4 * Hp, 4 * Wp, 3) and got.dtype == torch.uint8
        assert torch.equal(got, want)
        assert got.float().std() > 20                                # real content, both clamps exercised
    with pytest.raises(ValueError):
        ops.haar_unpatch_u8(torch.zeros(1, 2, 2, 64, device=DEV, dtype=torch.bfloat16))


def test_decode_u8_equals_decode_and_postprocess_end_to_end():
    from drb200 import ops
    vae, _ = _product_vae(vo.SMALL_VAE, seed=7)
    z = torch.randn(1, 16, 3, 4, 6, device=DEV, generator=gen(52)).bfloat16()
    for flag in (False, True):
        want = ops.postprocess_u8(vae.decode_scaled(z, 2.0)[0].contiguous(), flag)
        assert torch.equal(vae.decode_u8(z, 2.0, flag)[0], want)
