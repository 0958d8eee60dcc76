"""One representative tokenizer convolution at full size for profiling: (1,3,3) 512 -> 512 at 15 x 176 x 320 (decoder
level N0, 3.99e12 flop) and the (3,1,1) temporal convolution of the same tensor.  Prints time and TFLOP/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import ops

T, H, W, C = 15, 176, 320, 512
x = torch.randn(T, H, W, C, device="cuda").bfloat16()
ws = (torch.randn(C, 1, 3, 3, C, device="cuda") / (9 * C) ** 0.5).bfloat16()
wt = (torch.randn(C, 3, 1, 1, C, device="cuda") / (3 * C) ** 0.5).bfloat16()
b = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(T, 2, device="cuda", dtype=torch.float64)
for name, w, pad, taps in (("conv (1,3,3) 512->512", ws, 1, 9), ("conv (3,1,1) 512->512", wt, 0, 3)):
    ops.conv3d_cl(x, w, b, pad_h=pad, pad_w=pad, stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = int(os.environ.get("PROBE_ITERS", "5"))
    e0.record()
    for _ in range(n):
        ops.conv3d_cl(x, w, b, pad_h=pad, pad_w=pad, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * T * H * W * C * C * taps
    print(f"{name} @ {T}x{H}x{W}: {ms:.3f} ms, {fl / ms / 1e9:.0f} TFLOP/s")
