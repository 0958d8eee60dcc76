"""GPU unit parity of every sm_100a kernel behind the C ABI against the oracle's op-level functions (torch, same GPU).

bf16 GEMM / attention: relative L2 vs an fp32 reference <= 3e-3 (GEMM, bf16 output rounding ~2e-3 worst case) and
<= 5e-3 (attention; P is rounded to bf16 before P.V as in every flash kernel).  Elementwise kernels round where the
reference's chain of bf16 tensor ops rounds, so they are compared for (near) bit-equality: at most 1 bf16 ulp on a
small fraction of elements (transcendental / reduction-order differences), stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import dit_oracle as do
from oracle import sampler_oracle as so
from oracle.weights import DitDims
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def gen(seed=0):
    return torch.Generator(device=DEV).manual_seed(seed)


def assert_close_bf16(got, ref, max_ulp=1, min_exact=0.98):
    """|got - ref| <= max_ulp bf16 ulps of max(|ref|, max|ref| / 64): elements that are small through cancellation
    carry the absolute fp32 accumulation-order noise of the large terms they were summed from."""
    got, ref = got.float(), ref.float()
    tol = ref.abs().clamp_min(ref.abs().max() / 64) * (2.0 ** -7) * max_ulp     # 1 bf16 ulp <= 2^-7 |x|
    bad = (got - ref).abs() > tol
    assert not bad.any(), f"{int(bad.sum())} elements differ by more than {max_ulp} bf16 ulp"
    exact = (got == ref).float().mean().item()
    assert exact >= min_exact, f"only {exact:.4f} of the elements are bit-equal"


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (300, 520, 136), (48, 64, 256), (1000, 768, 616), (2048, 512, 2048),
                                   (3520, 4096, 256),    # a context-parallel shard (ragged last m-tile)
                                   (3520, 1152, 136)])
def test_gemm_store(cg, M, N, K):
    from drb200 import ops
    g = gen(1)
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    out = ops.gemm(a, w, cta_group=cg)
    ref = a.float() @ w.float().t()
    assert rel_l2(out, ref) <= 3e-3
    assert_close_bf16(out, ref.bfloat16(), max_ulp=1, min_exact=0.97)


@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_strided_views_and_epilogues(cg):
    from drb200 import ops
    g = gen(2)
    M, N, K = 640, 512, 384
    big = (torch.randn(M, 3 * K, device=DEV, generator=g) * 0.5).bfloat16()
    a = big[:, K:2 * K]                                               # row-strided view (like q/k/v of one buffer)
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    acc = (a.float() @ w.float().t()).bfloat16()
    # acc itself may differ by 1 bf16 ulp (fp32 accumulation order), which the epilogue math can amplify slightly
    assert_close_bf16(ops.gemm(a, w, epilogue=1, cta_group=cg), F.gelu(acc), max_ulp=2, min_exact=0.95)
    resid = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    gate = torch.randn(N, device=DEV, generator=g).bfloat16()
    ref = resid + gate[None, :] * acc
    x = resid.clone()
    ops.gemm(a, w, out=x, epilogue=2, resid=x, gate=gate, cta_group=cg)   # in place, as the DiT uses it
    assert_close_bf16(x, ref, max_ulp=4, min_exact=0.95)


def test_gemm_rejects_bad_arguments():
    from drb200 import ops
    a = torch.zeros(16, 12, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(8, 12, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.gemm(a, w)                                                # K not a multiple of 8
    with pytest.raises(ValueError):
        ops.gemm(a.float(), w)
    with pytest.raises(ValueError):
        ops.gemm(a.cpu(), w.cpu())
    a8 = torch.zeros(16, 16, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.gemm(a8, a8, epilogue=3)                                  # the QKV epilogue is internal to drb_gemm_qkv_norm_rope


def _sdpa_ref(q, k, v, H, dtype=torch.float32):
    S, Skv = q.shape[0], k.shape[0]
    q4 = q.to(dtype).reshape(S, H, 128).permute(1, 0, 2)[None]
    k4 = k.to(dtype).reshape(Skv, H, 128).permute(1, 0, 2)[None]
    v4 = v.to(dtype).reshape(Skv, H, 128).permute(1, 0, 2)[None]
    return F.scaled_dot_product_attention(q4, k4, v4)[0].permute(1, 0, 2).reshape(S, H * 128)


def _max_abs_logit(q, k, H):
    S, Skv = q.shape[0], k.shape[0]
    qh = q.float().reshape(S, H, 128).permute(1, 0, 2)
    kh = k.float().reshape(Skv, H, 128).permute(1, 0, 2)
    return (qh @ kh.transpose(1, 2)).abs().max().item() / 128 ** 0.5


@pytest.mark.parametrize("certified", [False, True], ids=["safe", "certified"])
@pytest.mark.parametrize("S,Skv,H", [(128, 128, 1), (512, 512, 4), (48, 48, 2), (1000, 1000, 2), (300, 700, 3), (4096, 4096, 2),
                                     (512, 1000, 2)])   # the last: K/V multicast pair (two query blocks) with a ragged key tile
def test_attention(S, Skv, H, certified):
    """both softmax flavours: the per-tile-max one (no certificate) and the max-free one under a logit bound"""
    from drb200 import ops
    g = gen(3)
    qkv = torch.randn(max(S, Skv), 3 * H * 128, device=DEV, generator=g).bfloat16()
    q, k, v = qkv[:S, :H * 128], qkv[:Skv, H * 128:2 * H * 128], qkv[:Skv, 2 * H * 128:]
    bound = None
    if certified:
        assert _max_abs_logit(q, k, H) <= 39.0
        bound = torch.tensor([39.0], device=DEV)
    out = ops.attention(q, k, v, H, max_abs_logit=bound)
    assert rel_l2(out, _sdpa_ref(q, k, v, H)) <= 5e-3
    # against the reference's own bf16 SDPA call (CleanGeneralDIT.py:192-197)
    assert rel_l2(out, _sdpa_ref(q, k, v, H, torch.bfloat16)) <= 8e-3


@pytest.mark.parametrize("certified", [False, True], ids=["safe", "certified"])
def test_attention_large_logits_rescale_path(certified):
    """row maxima that keep growing along kv force the O-rescale branch of either flavour"""
    from drb200 import ops
    g = gen(4)
    S, H = 1024, 1
    qs, k1 = (2.0, 3.0) if certified else (4.0, 6.0)
    q = torch.randn(S, 128, device=DEV, generator=g).bfloat16() * qs
    k = torch.randn(S, 128, device=DEV, generator=g).bfloat16()
    k = (k * torch.linspace(0.2, k1, S, device=DEV)[:, None]).bfloat16()      # later keys -> larger |logits|
    v = torch.randn(S, 128, device=DEV, generator=g).bfloat16()
    bound = None
    if certified:
        m = _max_abs_logit(q, k, H)
        assert 12.0 <= m <= 39.0, m          # far above tile 0's maximum (the reference must move), inside the certificate
        bound = torch.tensor([m], device=DEV)
    out = ops.attention(q, k, v, H, max_abs_logit=bound)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, _sdpa_ref(q, k, v, H)) <= 8e-3


@pytest.mark.parametrize("Sq,Skv,where", [(1024, 1024, 700), (512, 1000, 999), (300, 4096, 2500)])
def test_attention_sudden_outlier_logit(Sq, Skv, where):
    """One key deep in the sequence scores ~100 natural units above everything before it (2^144 in the log2 domain: a
    max-free softmax anchored at tile 0 would overflow).  Without a certificate the per-tile-max softmax must stay exact."""
    from drb200 import ops
    g = gen(11)
    H = 2
    q = (torch.randn(Sq, H * 128, device=DEV, generator=g) + 3.0).bfloat16()
    k = torch.randn(Skv, H * 128, device=DEV, generator=g).bfloat16()
    k[where] = 3.0                                                         # q.k / sqrt(128) ~ 128*3*3/11.3 = 102
    v = torch.randn(Skv, H * 128, device=DEV, generator=g).bfloat16()
    logits = (q.float()[:, :128] @ k.float()[:, :128].t()) / 128 ** 0.5
    assert (logits[:, where] - logits[:, :where].max(dim=1).values).min().item() > 60.0
    out = ops.attention(q, k, v, H)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, _sdpa_ref(q, k, v, H)) <= 8e-3
    # a certificate that the data violates must not be honoured blindly when it is above the kernel's limit
    out2 = ops.attention(q, k, v, H, max_abs_logit=torch.tensor([150.0], device=DEV))
    assert torch.equal(out, out2)


def test_qk_logit_bound_certifies_normed_heads():
    """drb_qk_logit_bound >= the largest |q.k|/sqrt(128) the per-head RMSNorm + RoPE can produce with these weights"""
    from drb200 import ops
    g = gen(12)
    L, S = 3, 256
    wq = (1.0 + 0.3 * torch.randn(L, 128, device=DEV, generator=g)).bfloat16()
    wk = (1.0 + 0.3 * torch.randn(L, 128, device=DEV, generator=g)).bfloat16()
    b = ops.qk_logit_bound(wq, wk)
    want = 128 ** 0.5 * wq.float().abs().amax(1) * wk.float().abs().amax(1) * 1.02
    assert torch.allclose(b, want, rtol=1e-5)
    ang = torch.rand(S, 128, device=DEV, generator=g) * 6.28
    cos, sin = ang.cos().bfloat16(), ang.sin().bfloat16()
    for l in range(L):
        # adversarial rows: all the energy of a head in the dimension with the largest weight
        qkv = torch.randn(S, 3 * 128, device=DEV, generator=g).bfloat16() * 0.01
        qkv[:, wq[l].float().abs().argmax()] = 50.0
        qkv[:, 128 + wk[l].float().abs().argmax()] = 50.0
        ops.qk_norm_rope(qkv, wq[l], wk[l], cos, sin, 1)
        assert _max_abs_logit(qkv[:, :128], qkv[:, 128:256], 1) <= b[l].item()
    wq[1, 5] = float("nan")
    assert torch.isinf(ops.qk_logit_bound(wq, wk)[1])


@pytest.mark.parametrize("D", [256, 512, 4096])
@pytest.mark.parametrize("with_add", [False, True])
def test_adaln_modulate(D, with_add):
    from drb200 import ops
    g = gen(5)
    rows = 77
    x = (torch.randn(rows, D, device=DEV, generator=g) * 3 + 0.5).bfloat16()
    shift = torch.randn(D, device=DEV, generator=g).bfloat16()
    scale = torch.randn(D, device=DEV, generator=g).bfloat16()
    gate = torch.randn(D, device=DEV, generator=g).bfloat16()
    vec = torch.randn(D, device=DEV, generator=g).bfloat16()
    xin = x.clone()
    if with_add:
        out = ops.adaln_modulate(xin, shift, scale, add_gate=gate, add_vec=vec)
        x_ref = x + gate[None] * vec[None]
        assert torch.equal(xin, x_ref)                                # residual stream updated in place, bit-exact
    else:
        out = ops.adaln_modulate(xin, shift, scale)
        x_ref = x
    ref = F.layer_norm(x_ref, (D,), eps=1e-6) * (1 + scale[None]) + shift[None]
    assert_close_bf16(out, ref, max_ulp=2, min_exact=0.97)


@pytest.mark.parametrize("H", [2, 32])
def test_qk_norm_rope(H):
    from drb200 import ops
    from drb200.CleanGeneralDIT import _RoPE3D
    g = gen(6)
    T, Hp, Wp = 2, 5, 7
    S, D = T * Hp * Wp, H * 128
    qkv = torch.randn(S, 3 * D, device=DEV, generator=g).bfloat16()
    wq = (1 + 0.1 * torch.randn(128, device=DEV, generator=g)).bfloat16()
    wk = (1 + 0.1 * torch.randn(128, device=DEV, generator=g)).bfloat16()
    dims = DitDims(model_channels=D, num_heads=H)
    ang = do.rope_angles(dims, T, Hp, Wp, torch.bfloat16, DEV)
    cos, sin = _RoPE3D(128).to(DEV).to(torch.bfloat16).tables(T, Hp, Wp, torch.bfloat16)
    assert torch.equal(cos, ang.cos()) and torch.equal(sin, ang.sin())   # host table == reference's bf16-angle quirk
    ref_q = do.apply_rope(do.rms_norm(qkv[:, :D].reshape(S, 1, H, 128), wq), ang).reshape(S, D)
    ref_k = do.apply_rope(do.rms_norm(qkv[:, D:2 * D].reshape(S, 1, H, 128), wk), ang).reshape(S, D)
    v_before = qkv[:, 2 * D:].clone()
    ops.qk_norm_rope(qkv, wq, wk, cos, sin, H)
    assert torch.equal(qkv[:, 2 * D:], v_before)
    assert_close_bf16(qkv[:, :D], ref_q, max_ulp=2, min_exact=0.97)
    assert_close_bf16(qkv[:, D:2 * D], ref_k, max_ulp=2, min_exact=0.97)


def test_gemv_and_sigma_embedding():
    from drb200 import ops
    g = gen(7)
    D = 512
    w = (torch.randn(3 * D, D, device=DEV, generator=g) / math.sqrt(D)).bfloat16()
    x = torch.randn(D, device=DEV, generator=g).bfloat16()
    add = torch.randn(3 * D, device=DEV, generator=g).bfloat16()
    assert_close_bf16(ops.gemv(w, x), F.linear(x[None], w)[0], max_ulp=1, min_exact=0.9)
    assert_close_bf16(ops.gemv(w, x, add=add, act=1), F.linear(F.silu(x)[None], w)[0] + add, max_ulp=1, min_exact=0.9)
    wb = (torch.randn(5, 3 * D, 256, device=DEV, generator=g) / 16).bfloat16()
    xb = torch.randn(5, 256, device=DEV, generator=g).bfloat16()
    out = torch.empty(5, 3 * D, device=DEV, dtype=torch.bfloat16)
    ops.gemv_batched(wb, xb, out, add=add)
    ref = torch.einsum("bnk,bk->bn", wb.float(), xb.float()).bfloat16() + add[None]
    assert_close_bf16(out, ref, max_ulp=1, min_exact=0.9)
    w_aff = (1 + 0.1 * torch.randn(D, device=DEV, generator=g)).bfloat16()
    for s in (80.0, 3.17, 0.02):
        sig = torch.tensor([s], device=DEV)
        e, emb = torch.empty(D, device=DEV, dtype=torch.bfloat16), torch.empty(D, device=DEV, dtype=torch.bfloat16)
        ops.sigma_embedding(sig, w_aff, e, emb)
        ref_e = do.sigma_embedding(sig.bfloat16(), D)[0]
        assert_close_bf16(e, ref_e, max_ulp=1, min_exact=0.9)
        assert_close_bf16(emb, do.rms_norm(ref_e[None], w_aff)[0], max_ulp=1, min_exact=0.9)


def test_patchify_unpatchify_euler():
    from drb200 import ops
    g = gen(8)
    C, T, H, W, Cc = 16, 2, 8, 12, 16
    S = T * (H // 2) * (W // 2)
    x = (torch.randn(C, T, H, W, device=DEV, generator=g) * 20).bfloat16()
    cond = torch.randn(Cc, T, H, W, device=DEV, generator=g).bfloat16()
    sig, nxt = torch.tensor([7.3], device=DEV), torch.tensor([2.9], device=DEV)
    kdim = (C + Cc + 1) * 4
    tok = torch.full((S, kdim + 4), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.patchify_condition(cond, tok, C, T, H, W, ones_channel=C + Cc, zero_from=kdim)
    ops.scale_patchify(x, sig, tok)
    xin = so.scale_model_input(x[None], sig[0])
    full = torch.cat([xin, cond[None], torch.ones(1, 1, T, H, W, device=DEV, dtype=torch.bfloat16)], dim=1)
    ref = do.patchify(full, 2).reshape(S, kdim)
    assert torch.equal(tok[:, :kdim], ref) and (tok[:, kdim:] == 0).all()
    y = torch.randn(S, 4 * C, device=DEV, generator=g).bfloat16()
    yu = torch.randn(S, 4 * C, device=DEV, generator=g).bfloat16()
    f_ref = do.unpatchify(y, 1, T, H // 2, W // 2, 2, C)
    f_out = torch.empty(C, T, H, W, device=DEV, dtype=torch.bfloat16)
    ops.unpatchify_euler(y, None, 0.0, None, None, None, None, f_out=f_out)
    assert torch.equal(f_out[None], f_ref)
    x_next = torch.empty_like(x)
    ops.unpatchify_euler(y, None, 0.0, sig, nxt, x, x_next)
    assert_close_bf16(x_next[None], so.euler_step(f_ref, sig[0], nxt[0], x[None]), max_ulp=1, min_exact=0.98)
    fu = do.unpatchify(yu, 1, T, H // 2, W // 2, 2, C)
    f_cfg = f_ref + 2.0 * (f_ref - fu)
    ops.unpatchify_euler(y, yu, 2.0, sig, nxt, x, x_next, f_out=f_out)
    assert torch.equal(f_out[None], f_cfg)
    assert_close_bf16(x_next[None], so.euler_step(f_cfg, sig[0], nxt[0], x[None]), max_ulp=1, min_exact=0.98)
    last = torch.zeros(1, device=DEV)
    ops.unpatchify_euler(y, None, 0.0, sig, last, x, x_next)          # sigma_next = 0: x_next = denoised
    assert_close_bf16(x_next[None], so.euler_step(f_ref, sig[0], last[0], x[None]), max_ulp=1, min_exact=0.98)


@pytest.mark.parametrize("normalize", [False, True])
def test_postprocess_u8(normalize):
    from drb200 import ops
    g = gen(9)
    v = (torch.randn(1, 3, 3, 16, 24, device=DEV, generator=g) * 0.8).bfloat16()
    got = ops.postprocess_u8(v[0], normalize).cpu().numpy()
    ref = so.postprocess(v, normalize)[0]
    diff = abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= (2 if normalize else 0)      # plain path is bit-exact; the normal blend may move 1 bf16 ulp
    assert (diff == 0).mean() >= 0.97


@pytest.mark.parametrize("M,D,K", [(300, 256, 256), (1000, 512, 512), (77, 256, 136), (3520, 1024, 256)])
def test_fused_qkv_gemm_norm_rope_matches_the_unfused_kernels(M, D, K):
    """drb_gemm_qkv_norm_rope == drb_gemm_bf16 followed by drb_qk_norm_rope, up to the summation order of the RMS"""
    from drb200 import ops
    g = gen(11)
    H = D // 128
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    w = (torch.randn(3 * D, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    wq = (1 + 0.1 * torch.randn(128, device=DEV, generator=g)).bfloat16()
    wk = (1 + 0.1 * torch.randn(128, device=DEV, generator=g)).bfloat16()
    ang = (torch.rand(M, 128, device=DEV, generator=g) * 6.28).bfloat16()
    cos, sin = ang.cos().bfloat16().contiguous(), ang.sin().bfloat16().contiguous()
    ref = ops.gemm(a, w)
    ops.qk_norm_rope(ref, wq, wk, cos, sin, H)
    got = ops.qkv_gemm_norm_rope(a, w, wq, wk, cos, sin)
    assert torch.equal(got[:, 2 * D:], ref[:, 2 * D:])                      # v: plain projection
    # the RMS is summed in a different order (one thread vs a warp tree): a flipped bf16 rounding of the normalised value
    # shows up as 1 ulp of a RoPE product, i.e. possibly several ulp of a sum that cancels — gate on the absolute scale
    qk_got, qk_ref = got[:, :2 * D].float(), ref[:, :2 * D].float()
    assert (qk_got == qk_ref).float().mean() >= 0.995
    assert (qk_got - qk_ref).abs().max() <= 2.0 ** -7 * qk_ref.abs().max()
    assert rel_l2(qk_got, qk_ref) <= 1e-3
    # context-parallel form: rows scattered to the (virtual) head owners
    world, S_tot, row0 = 2, M + 50, 13
    Dp = D // world
    bufs = [torch.zeros(S_tot, 3 * Dp, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
    ops.qkv_gemm_norm_rope(a, w, wq, wk, cos, sin, peer_ptrs=[b.data_ptr() for b in bufs], peer_ld=3 * Dp, row0=row0)
    for r, b in enumerate(bufs):
        for sect in range(3):
            assert torch.equal(b[row0:row0 + M, sect * Dp:(sect + 1) * Dp], got[:, sect * D + r * Dp: sect * D + (r + 1) * Dp])
        assert not b[:row0].any() and not b[row0 + M:].any()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_kernels_are_configured_per_device():
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) belongs to a device's context: a process that drives a second GPU
    must get its kernels configured there too (the round-1 library kept one process-wide flag and failed on device 1)"""
    from drb200 import ops
    outs = []
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            dev = f"cuda:{d}"
            g = torch.Generator(device=dev).manual_seed(1)
            a = torch.randn(300, 256, device=dev, generator=g).bfloat16()
            w = torch.randn(512, 256, device=dev, generator=g).bfloat16()
            qkv = torch.randn(300, 3 * 256, device=dev, generator=g).bfloat16()
            x = torch.randn(2, 16, 16, 64, device=dev, generator=g).bfloat16()
            wc = (torch.randn(64, 1, 3, 3, 64, device=dev, generator=g) * 0.05).bfloat16()
            b = torch.zeros(64, device=dev, dtype=torch.bfloat16)
            r = (ops.gemm(a, w), ops.attention(qkv[:, :256], qkv[:, 256:512], qkv[:, 512:], 2),
                 ops.conv3d_cl(x, wc, b, pad_h=1, pad_w=1), ops.haar_unpatch(torch.randn(1, 2, 3, 192, device=dev, generator=g).bfloat16()))
            torch.cuda.synchronize(d)
            outs.append([t.cpu() for t in r])
    for p, q in zip(outs[0], outs[2]):
        assert torch.equal(p, q)
    for p, q in zip(outs[0], outs[1]):        # same seeds, same kernels, other device
        assert torch.equal(p, q)
