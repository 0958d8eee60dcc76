"""The "existing Blackwell path" (second baseline): what the reference executes on one B200 through stock
PyTorch — cuBLASLt linears, F.scaled_dot_product_attention, unfused elementwise tensor ops — for ONE FA-CA-MLP block of
the 7B GeneralDIT at S = 28 160 tokens in bf16, written with plain torch calls in the reference's operator order
(CleanGeneralDIT.py:268-306, :442-462, :492-517; the head-flatten patch of SURVEY.md D1 applied).  Scaled x28 for a
forward.  Not part of the product and independent of oracle/: it only measures the library path on the same box.

    python tools/torch_gpu_baseline.py [tokens] [min_seconds]"""
import sys
import time

import torch
import torch.nn.functional as F

D, H, HID, R = 4096, 32, 16384, 256


def measure(S: int = 28160, min_seconds: float = 3.0, min_repeats: int = 5) -> dict:
    """Run the block back to back for at least `min_seconds` (sustained, power-capped clocks like the product's step) and
    return the bench.py `gpu_baseline` record."""
    dev, dt = "cuda", torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)

    def w(n, k):
        return ((torch.rand(n, k, device=dev, generator=g) * 2 - 1) / k ** 0.5).to(dt)

    W = {n: w(D, D) for n in ("q", "k", "v", "o", "cq", "co")}
    W.update(ck=w(D, 1024), cv=w(D, 1024), l1=w(HID, D), l2=w(D, HID))
    MOD = [(w(R, D), w(3 * D, R)) for _ in range(3)]
    qn, kn, cqn, ckn = (torch.ones(128, device=dev, dtype=dt) for _ in range(4))
    x = torch.randn(S, 1, D, device=dev, generator=g).to(dt)
    emb = torch.randn(1, D, device=dev, generator=g).to(dt)
    lora = (0.1 * torch.randn(1, 3 * D, device=dev, generator=g)).to(dt)
    ctx = torch.randn(1, 1, 1024, device=dev, generator=g).to(dt)
    ang = torch.randn(S, 1, 1, 128, device=dev, generator=g).to(dt)

    def rmsnorm(t, wt):
        tf = t.float()
        return (tf * torch.rsqrt(tf.pow(2).mean(-1, keepdim=True) + 1e-6)).type_as(t) * wt

    def rope(t, freqs):
        cos, sin = torch.cos(freqs).to(t.dtype), torch.sin(freqs).to(t.dtype)
        t1, t2 = t.chunk(2, dim=-1)
        return t * cos + torch.cat((-t2, t1), dim=-1) * sin

    def attn(xq, ctxt, wq, wk, wv, wo, nq, nk, use_rope):
        q = rmsnorm(F.linear(xq, wq).view(xq.shape[0], 1, H, 128), nq)
        k = rmsnorm(F.linear(ctxt, wk).view(ctxt.shape[0], 1, H, 128), nk)
        v = F.linear(ctxt, wv).view(ctxt.shape[0], 1, H, 128)
        if use_rope:
            q, k = rope(q, ang), rope(k, ang)
        o = F.scaled_dot_product_attention(q.permute(1, 2, 0, 3), k.permute(1, 2, 0, 3), v.permute(1, 2, 0, 3))
        return F.linear(o.permute(2, 0, 1, 3).flatten(2), wo)

    def sub_block(xc, j, fn):
        a, b = MOD[j]
        m = F.linear(F.linear(F.silu(emb), a), b) + lora
        shift, scale, gate = m.chunk(3, dim=-1)
        xm = F.layer_norm(xc, (D,), eps=1e-6) * (1 + scale) + shift
        return xc + gate * fn(xm)

    @torch.no_grad()
    def block(xc):
        xc = sub_block(xc, 0, lambda t: attn(t, t, W["q"], W["k"], W["v"], W["o"], qn, kn, True))
        xc = sub_block(xc, 1, lambda t: attn(t, ctx, W["cq"], W["ck"], W["cv"], W["co"], cqn, ckn, False))
        return sub_block(xc, 2, lambda t: F.linear(F.gelu(F.linear(t, W["l1"])), W["l2"]))

    for _ in range(2):
        block(x)
    torch.cuda.synchronize()
    # sustained: keep the GPU busy for min_seconds, time the second half only (the clocks have settled under the power cap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    block(x)
    e1.record()
    torch.cuda.synchronize()
    rep = max(min_repeats, int(min_seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)) + 1)
    for _ in range(rep // 2):
        block(x)
    n = rep - rep // 2
    e0.record()
    for _ in range(n):
        block(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nec = 8 * S * D * D + 4 * S * S * D + 16 * S * D * D
    return {"value": 1000.0 / (28 * ms), "unit": "steps/s", "ms_per_block": ms, "blocks_run": rep + 3, "blocks_timed": n,
            "tflops_necessary_work": nec / ms / 1e9, "tokens": S,
            "what": f"stock PyTorch {torch.__version__} (cuBLASLt linears, F.scaled_dot_product_attention, unfused elementwise ops) "
                    "running the reference's operator sequence for ONE FA-CA-MLP block of the 7B GeneralDIT (CleanGeneralDIT.py:268-306, "
                    ":442-462, :492-517, head-flatten patch applied), back to back for >= 3 s on this GPU, x28 blocks; excludes the "
                    "embed / final layers and the sampler glue, so it flatters the baseline slightly"}


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 28160
    r = measure(S, float(sys.argv[2]) if len(sys.argv) > 2 else 3.0)
    print(f"torch {torch.__version__} on {torch.cuda.get_device_name()}: one FA-CA-MLP block at S={S}: {r['ms_per_block']:.2f} ms "
          f"-> 28-block forward {28 * r['ms_per_block']:.0f} ms = {r['value']:.3f} denoise steps/s; "
          f"{r['tflops_necessary_work']:.0f} TFLOP/s of necessary work per block ({r['blocks_timed']} blocks timed after "
          f"{r['blocks_run'] - r['blocks_timed']} warm-up blocks)")
