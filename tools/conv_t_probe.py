"""The tokenizer's temporal (3,1,1) convolutions at full size, with and without the two epilogue extras they carry in the
network (skip term, GroupNorm sums): which part bounds them?   python tools/conv_t_probe.py [C] [T H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import _lib, ops

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T, H, W = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (15, 176, 320)
x = torch.randn(T, H, W, C, device="cuda").bfloat16()
res = torch.randn(T, H, W, C, device="cuda").bfloat16()
wt = (torch.randn(C, 3, 1, 1, C, device="cuda") / (3 * C) ** 0.5).bfloat16()
b = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(T, 2, device="cuda", dtype=torch.float64)
out = torch.empty_like(x)
fl = 2.0 * T * H * W * C * C * 3
n = int(os.environ.get("PROBE_ITERS", "5"))
for name, kw in (("skip + sums", dict(resid=res, resid_mode=_lib.RES_SAME, stats=stats)), ("sums only", dict(stats=stats)),
                 ("skip only", dict(resid=res, resid_mode=_lib.RES_SAME)), ("plain", dict())):
    ops.conv3d_cl(x, wt, b, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.conv3d_cl(x, wt, b, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"conv (3,1,1) {C}->{C} @ {T}x{H}x{W}, {name:12s}: {ms:.3f} ms, {fl / ms / 1e9:.0f} TFLOP/s, "
          f"{(2 + ('resid' in kw)) * x.numel() * 2 / ms / 1e6:.0f} GB/s algorithmic")
