"""Per-kernel device time of one tokenizer encode + decode (CUDA events around every C-ABI call, aggregated by kernel and, for
the convolutions, by shape) — where the tokenizer's time goes and at what TFLOP/s each convolution shape runs.

    python tools/vae_kernel_probe.py [frames height width]   (default 57 704 1280)"""
import collections
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import _lib
from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE

T, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (57, 704, 1280)
model = AutoencoderKLCosmos()
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    for n, p in model.named_parameters():
        if n.endswith("bias"):
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
vae = CleanVAE(model=model)
vae.to("cuda")
vae.reset_dtype(torch.bfloat16)
x = (torch.rand(1, 3, T, H, W, device="cuda") * 2 - 1).bfloat16()
for _ in range(2):
    z = vae.encode(x)
    y = vae.decode(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    y = vae.decode(vae.encode(x))
e1.record()
torch.cuda.synchronize()
print(f"encode + decode of {T}x{H}x{W}: {e0.elapsed_time(e1) / 3:.2f} ms (plain timing)")

real_call = _lib.call
recs = []


def timed_call(name, *args):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    key, flop = name.replace("drb_", ""), 0.0
    if name == "drb_conv3d_cl":
        c = args[0]
        key = (f"conv {c.kt}x{c.kh}x{c.kw} {c.Cin:4d}->{c.Cout:4d} @ {c.T_out}x{c.H_out}x{c.W_out}"
               f"{' s2' if c.stride_hw == 2 else ''}{' tm%d' % c.tmode if c.tmode else ''}{' sub' if c.out_scale == 2 else ''}")
        flop = 2.0 * c.T_out * c.H_out * c.W_out * c.Cout * c.kt * c.kh * c.kw * c.Cin
    elif name == "drb_spatial_attention_d512":
        flop = 4.0 * args[4] * args[5] * args[5] * 512
    elif name == "drb_gemm_bf16":
        key = f"gemm M={args[6]} N={args[7]} K={args[8]}"
        flop = 2.0 * args[6] * args[7] * args[8]
    a.record()
    real_call(name, *args)
    b.record()
    recs.append((key, flop, a, b))


_lib.call = timed_call
for phase, fn in (("encode", lambda: vae.encode(x)), ("decode", lambda: vae.decode(z))):
    recs.clear()
    fn()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for key, flop, a, b in recs:
        t, c, f = agg.get(key, (0.0, 0, 0.0))
        agg[key] = (t + a.elapsed_time(b), c + 1, f + flop)
    total = sum(t for t, _, _ in agg.values())
    print(f"\n=== {phase}: {total:.2f} ms summed over {len(recs)} calls")
    for key, (t, c, f) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        tf = f" {f / t / 1e9:7.0f} TFLOP/s" if f else ""
        print(f"  {key:58s} x{c:3d} {t:8.3f} ms {100 * t / total:5.1f}%{tf}")
_lib.call = real_call
