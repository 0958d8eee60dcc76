"""Times the B200 tokenizer (CleanVAE encode / decode) on a full-size clip with random-init weights; product path only.
usage: python tools/vae_probe.py [frames height width]   (default 57 704 1280)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import _lib
from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE

T, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (57, 704, 1280)
ENC_FLOP = {57: 1.764e13, 121: 3.566e13}
DEC_FLOP = {57: 3.014e13, 121: 6.128e13}
torch.manual_seed(0)
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


def timed(fn, n=int(os.environ.get("PROBE_ITERS", "3"))):
    fn()
    torch.cuda.synchronize()
    l0 = _lib.LAUNCHES
    best = 1e9
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn()
        e1.record()
        t_host = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        print(f"   iter {i}: device {e0.elapsed_time(e1):.1f} ms, host enqueue {t_host:.1f} ms")
        best = min(best, e0.elapsed_time(e1))
    return best, (_lib.LAUNCHES - l0) // n, out


scale = H * W / (704 * 1280)
ms, nl, z = timed(lambda: vae.encode(x))
print(f"encode {T}x{H}x{W}: {ms:.1f} ms, {nl} launches, latent {tuple(z.shape)}, finite={bool(torch.isfinite(z.float()).all())}", end="")
if T in ENC_FLOP:
    print(f", {ENC_FLOP[T] * scale / ms / 1e9:.0f} TFLOP/s algorithmic")
else:
    print()
ms, nl, y = timed(lambda: vae.decode(z))
print(f"decode: {ms:.1f} ms, {nl} launches, video {tuple(y.shape)}, finite={bool(torch.isfinite(y.float()).all())}", end="")
if T in DEC_FLOP:
    print(f", {DEC_FLOP[T] * scale / ms / 1e9:.0f} TFLOP/s algorithmic")
else:
    print()
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
