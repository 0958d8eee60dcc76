"""GroupNorm(+SiLU) apply at the tokenizer's largest shape, for profiling: x [15,176,320,256] bf16 (432 MB in, 432 MB out)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import ops

T, H, W, C = 15, 176, 320, 256
x = torch.randn(T, H, W, C, device="cuda").bfloat16()
gamma = torch.ones(C, device="cuda", dtype=torch.bfloat16)
beta = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
stats = ops.frame_stats(x)
out = torch.empty_like(x)
for silu in (True, False):
    ops.groupnorm_apply(x, stats, gamma, beta, silu, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = int(os.environ.get("PROBE_ITERS", "10"))
    e0.record()
    for _ in range(n):
        ops.groupnorm_apply(x, stats, gamma, beta, silu, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"groupnorm_apply silu={silu}: {ms:.3f} ms, {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")
y = torch.empty_like(x)
y.copy_(x)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    y.copy_(x)
e1.record()
torch.cuda.synchronize()
print(f"torch copy of the same tensor: {e0.elapsed_time(e1) / 10:.3f} ms, {2 * x.numel() * 2 / (e0.elapsed_time(e1) / 10) / 1e6:.0f} GB/s")
