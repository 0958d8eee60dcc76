"""A/B of the two softmax flavours of drb_attention_bf16* at the bench shape (S = 28 160, 32 heads): certified (max-free, lazy
reference) vs safe (per-tile maximum).  Burst timing, L2-cold inputs (230 MB each), CUDA events.

    python tools/attn_ab.py [S] [H] [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 28160
H = int(sys.argv[2]) if len(sys.argv) > 2 else 32
REP = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(S, 3 * H * 128, device=dev, generator=g).bfloat16()
q, k, v = qkv[:, :H * 128], qkv[:, H * 128:2 * H * 128], qkv[:, 2 * H * 128:]
out = torch.empty(S, H * 128, device=dev, dtype=torch.bfloat16)
bound = torch.tensor([30.0], device=dev)
flops = 4.0 * S * S * 128 * H
res = {}
for name, b in (("certified", bound), ("safe", None), ("certified", bound), ("safe", None)):
    for _ in range(3):
        ops.attention(q, k, v, H, out=out, max_abs_logit=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REP):
        ops.attention(q, k, v, H, out=out, max_abs_logit=b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REP
    res.setdefault(name, []).append(ms)
    print(f"{name:10s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
a = ops.attention(q, k, v, H, max_abs_logit=bound)
b = ops.attention(q, k, v, H)
print(f"certified vs safe rel-L2 {((a.float() - b.float()).norm() / b.float().norm()).item():.2e}")
