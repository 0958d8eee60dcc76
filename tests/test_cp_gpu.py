"""GPU parity of the context-parallel GeneralDIT path (csrc/cp.cu, attention epilogue with peer destinations).

Single GPU: P virtual ranks in one process (context_parallel.EmulatedGroup) run the same kernels with the same pointer
tables; every stage is issued for all ranks before the next, which replaces the device barrier.  Because every GEMM /
AdaLN row and every attention row is computed exactly as on one GPU (same tiles, same order), the result must be
BIT-IDENTICAL to the single-GPU forward.  Two or more GPUs: tools/cp_check.py under torchrun (real CUDA IPC buffers,
P2P stores over NVLink, device barrier); skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

from oracle.weights import TINY_FORWARD, TINY_INVERSE
from tests.util import build_product_model

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ring_blocks_merge_to_the_full_attention():
    """drb_attention_bf16_ring over K/V blocks of uneven length == one attention over all keys (fp32 running state)"""
    import torch.nn.functional as F
    from drb200 import ops
    from tests.util import rel_l2
    g = torch.Generator(device=DEV).manual_seed(9)
    Sq, H = 300, 2
    blocks = [130, 257, 64, 500]                                  # ragged tiles inside blocks; logits grow along the sequence
    Skv = sum(blocks)
    q = torch.randn(Sq, H * 128, device=DEV, generator=g).bfloat16()
    k = (torch.randn(Skv, H * 128, device=DEV, generator=g) * torch.linspace(0.3, 3.0, Skv, device=DEV)[:, None]).bfloat16()
    v = torch.randn(Skv, H * 128, device=DEV, generator=g).bfloat16()
    st_o = torch.empty(Sq, H * 128, device=DEV, dtype=torch.float32)
    st_ml = torch.empty(Sq, H, 2, device=DEV, dtype=torch.float32)
    out = torch.empty(Sq, H * 128, device=DEV, dtype=torch.bfloat16)
    start = 0
    for i, n in enumerate(blocks):
        ops.attention_ring_block(q, k[start:start + n], v[start:start + n], H, st_o, st_ml, first=(i == 0), last=(i == len(blocks) - 1),
                                 out=out)
        start += n
    ref = F.scaled_dot_product_attention(*(t.float().reshape(-1, H, 128).permute(1, 0, 2)[None] for t in (q, k, v)))[0]
    ref = ref.permute(1, 0, 2).reshape(Sq, H * 128)
    assert rel_l2(out, ref) <= 5e-3
    assert rel_l2(out, ops.attention(q, k, v, H)) <= 4e-3


@pytest.mark.parametrize("world", [2, 4])
def test_emulated_ring_matches_one_gpu(world):
    """ring mode (heads whole, K/V blocks pulled from the other ranks' buffers): same forward up to the order in which the
    key blocks enter the softmax"""
    from drb200 import ops
    from drb200.context_parallel import EmulatedGroup, shard_frames
    from tests.util import rel_l2
    dims = TINY_INVERSE
    model, _ = build_product_model(dims, "inverse", seed=3)
    net = model.net
    T, H, W = 4, 12, 20
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(1, 16, T, H, W, device=DEV, generator=g).bfloat16()
    cond = (torch.randn(1, 16, T, H, W, device=DEV, generator=g) * 0.5).bfloat16()
    ci = torch.full((1, 1), 2, dtype=torch.long, device=DEV)
    sigma = torch.tensor(1.26, device=DEV)
    with torch.no_grad():
        ref = net(x=x, timesteps=sigma, latent_condition=cond, context_index=ci)
        group = EmulatedGroup(world, mode="ring")
        wss, outs = [], []
        for cp in group.ranks:
            t0, t1 = shard_frames(T, cp.rank, world)
            ws = net._workspace(t1 - t0, H, W, x.device, cp)
            ws["sigma"].copy_(sigma.reshape(1))
            net.modulation(ws, ws["sigma"])
            net.prepare_condition(ws, cond[:, :, t0:t1], t1 - t0, H, W)
            ops.patchify_condition(x[0, :, t0:t1].contiguous(), ws["tok"], 0, t1 - t0, H, W)
            use_ca = net.prepare_context(ws, net.context_token(ci))
            net.stage_embed(ws)
            wss.append(ws)
        for i in range(net.num_blocks):
            for ws in wss:
                net.stage_pre_attention(ws, i)
            torch.cuda.synchronize()            # stands in for the device barrier (the ring uses a second stream)
            for ws in wss:
                net.stage_attention(ws, i)
            torch.cuda.synchronize()
            for ws in wss:
                net.stage_post_attention(ws, i, use_ca)
        for cp, ws in zip(group.ranks, wss):
            t0, t1 = shard_frames(T, cp.rank, world)
            out = torch.empty((16, t1 - t0, H, W), device=DEV, dtype=torch.bfloat16)
            ops.unpatchify_euler(net.stage_final(ws), None, 0.0, None, None, None, None, f_out=out)
            outs.append(out)
    got = torch.cat(outs, dim=1).unsqueeze(0)
    err = rel_l2(got, ref)
    print(f"\nring P={world}: F rel-L2 vs one GPU {err:.3e}")
    assert err <= 6e-3


@pytest.mark.parametrize("dims,mt", [(TINY_INVERSE, "inverse"), (TINY_FORWARD, "forward")])
@pytest.mark.parametrize("world", [2, 4])
def test_emulated_ranks_are_bit_identical_to_one_gpu(dims, mt, world):
    from drb200 import ops
    from drb200.context_parallel import EmulatedGroup, shard_frames
    model, _ = build_product_model(dims, mt, seed=3)
    net = model.net
    T, H, W = 4, 12, 20                                   # S = 240 tokens, 60 per rank at P = 4: ragged attention tiles
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(1, 16, T, H, W, device=DEV, generator=g).bfloat16()
    cond = (torch.randn(1, dims.additional_concat_ch, T, H, W, device=DEV, generator=g) * 0.5).bfloat16()
    ci = torch.full((1, 1), 2, dtype=torch.long, device=DEV)
    sigma = torch.tensor(1.26, device=DEV)
    with torch.no_grad():
        ref = net(x=x, timesteps=sigma, latent_condition=cond, context_index=ci)
        group = EmulatedGroup(world)
        wss, outs = [], []
        for cp in group.ranks:
            t0, t1 = shard_frames(T, cp.rank, world)
            ws = net._workspace(t1 - t0, H, W, x.device, cp)
            ws["sigma"].copy_(sigma.reshape(1))
            net.modulation(ws, ws["sigma"])
            net.prepare_condition(ws, cond[:, :, t0:t1], t1 - t0, H, W)
            ops.patchify_condition(x[0, :, t0:t1].contiguous(), ws["tok"], 0, t1 - t0, H, W)
            use_ca = net.prepare_context(ws, net.context_token(ci))
            net.stage_embed(ws)
            wss.append(ws)
        for i in range(net.num_blocks):
            for ws in wss:
                net.stage_pre_attention(ws, i)
            for ws in wss:
                net.stage_attention(ws, i)
            for ws in wss:
                net.stage_post_attention(ws, i, use_ca)
        for cp, ws in zip(group.ranks, wss):
            t0, t1 = shard_frames(T, cp.rank, world)
            out = torch.empty((16, t1 - t0, H, W), device=DEV, dtype=torch.bfloat16)
            ops.unpatchify_euler(net.stage_final(ws), None, 0.0, None, None, None, None, f_out=out)
            outs.append(out)
    got = torch.cat(outs, dim=1).unsqueeze(0)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("world", [2, 4])
def test_emulated_ranks_with_batched_passes_are_bit_identical_to_one_gpu(world):
    """B = 3 sequences (G-buffer passes with different context vectors) stacked along the token rows of every rank: the fused
    QKV epilogue scatters (sequence, head) pairs, the attention runs B*H/P problems per rank and scatters rows back.  S/P = 96
    or 48 rows per sequence: 128-row GEMM tiles and 32-row epilogue warps straddle sequence boundaries."""
    from drb200 import ops
    from drb200.context_parallel import EmulatedGroup, shard_frames
    dims = TINY_INVERSE
    model, _ = build_product_model(dims, "inverse", seed=3)
    net = model.net
    T, H, W, B = 4, 12, 16, 3                             # S = 192 tokens per sequence
    g = torch.Generator(device=DEV).manual_seed(5)
    xs = torch.randn(B, 16, T, H, W, device=DEV, generator=g).bfloat16()
    cond = (torch.randn(1, 16, T, H, W, device=DEV, generator=g) * 0.5).bfloat16()
    cis = [torch.full((1, 1), k, dtype=torch.long, device=DEV) for k in (1, 4, 2)]
    sigma = torch.tensor(1.26, device=DEV)
    with torch.no_grad():
        refs = [net(x=xs[b:b + 1], timesteps=sigma, latent_condition=cond, context_index=cis[b]) for b in range(B)]
        group = EmulatedGroup(world)
        wss = []
        for cp in group.ranks:
            t0, t1 = shard_frames(T, cp.rank, world)
            ws = net._workspace(t1 - t0, H, W, xs.device, cp, batch=B)
            ws["sigma"].copy_(sigma.reshape(1))
            net.modulation(ws, ws["sigma"])
            for b in range(B):
                net.prepare_condition(ws, cond[:, :, t0:t1], t1 - t0, H, W, b)
                ops.patchify_condition(xs[b, :, t0:t1].contiguous(), net._rows(ws, "tok", b), 0, t1 - t0, H, W)
                use_ca = net.prepare_context(ws, net.context_token(cis[b]), b)
            net.stage_embed(ws)
            wss.append(ws)
        for i in range(net.num_blocks):
            for ws in wss:
                net.stage_pre_attention(ws, i)
            for ws in wss:
                net.stage_attention(ws, i)
            for ws in wss:
                net.stage_post_attention(ws, i, use_ca)
        outs = [[None] * world for _ in range(B)]
        for cp, ws in zip(group.ranks, wss):
            t0, t1 = shard_frames(T, cp.rank, world)
            y = net.stage_final(ws)
            for b in range(B):
                out = torch.empty((16, t1 - t0, H, W), device=DEV, dtype=torch.bfloat16)
                ops.unpatchify_euler(y[b * ws["S"]:(b + 1) * ws["S"]], None, 0.0, None, None, None, None, f_out=out)
                outs[b][cp.rank] = out
    for b in range(B):
        assert torch.equal(torch.cat(outs[b], dim=1).unsqueeze(0), refs[b]), f"sequence {b}"


def test_cp_rejects_uneven_splits():
    from drb200.context_parallel import EmulatedGroup
    model, _ = build_product_model(TINY_INVERSE, "inverse", seed=3)
    with pytest.raises(ValueError):
        model.net.enable_context_parallel(EmulatedGroup(3).ranks[0])      # 4 heads over 3 ranks


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_two_gpu_context_parallel_matches_one_gpu():
    n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                        "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tools", "cp_check.py"), "--tiny"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "CP_CHECK_OK" in r.stdout
