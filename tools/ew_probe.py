"""Timing of the HBM-bound DiT kernels at the bench shape (S = 28160, D = 4096): achieved GB/s against algorithmic bytes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import ops

S, D = 28160, 4096
dev = "cuda"
x = torch.randn(S, D, device=dev).bfloat16()
out = torch.empty_like(x)
shift, scale, gate, vec = (torch.randn(D, device=dev).bfloat16() for _ in range(4))
big = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)   # L2 flush


def timed(fn, n=20):
    fn()
    ts = []
    for _ in range(n):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


ms = timed(lambda: ops.adaln_modulate(x, shift, scale, out=out))
print(f"adaln_modulate: {ms * 1e3:.1f} us, {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s (read x + write xm)")
ms = timed(lambda: ops.adaln_modulate(x, shift, scale, out=out, add_gate=gate, add_vec=vec))
print(f"adaln_modulate + CA residual: {ms * 1e3:.1f} us, {3 * x.numel() * 2 / ms / 1e6:.0f} GB/s (read x, write x and xm)")
ms = timed(lambda: out.copy_(x))
print(f"torch copy (reference point): {ms * 1e3:.1f} us, {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")
