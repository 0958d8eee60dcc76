"""The fused dim-512 spatial attention at the tokenizer's size (8 latent frames of 88x160 = 14 080 tokens), for profiling."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import ops

T, H, W = 8, 88, 160
qkv = torch.randn(T, H, W, 1536, device="cuda").bfloat16()
out = torch.empty(T, H, W, 512, device="cuda", dtype=torch.bfloat16)
n = int(os.environ.get("PROBE_ITERS", "5"))
ops.spatial_attention_d512(qkv, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    ops.spatial_attention_d512(qkv, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 4.0 * T * (H * W) ** 2 * 512
print(f"spatial attention d512, {T} frames x {H * W} tokens: {ms:.3f} ms, {fl / ms / 1e9:.0f} TFLOP/s algorithmic, {1.5 * fl / ms / 1e9:.0f} executed")
