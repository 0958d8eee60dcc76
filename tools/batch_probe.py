"""Per-kernel device time of one 7B denoise iteration with B passes batched along the token rows (single GPU), from CUDA
events recorded around every C-ABI call — answers "does a kernel get slower per row when M grows 5x?".

    python tools/batch_probe.py [B ...]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import _lib
from drb200 import diffusion_renderer_config as cfgm
from drb200.model_diffusion_renderer import CleanDiffusionRendererModel

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = cfgm.get_inverse_renderer_config(704, 1280, 57)
cfg["model_type"] = "inverse"
with torch.device("meta"):
    model = CleanDiffusionRendererModel(cfg)
model = model.to_empty(device=dev).to(torch.bfloat16)
net = model.net.init_weights_(seed=0)
net._ensure_packed()
T, H, W = 8, 88, 160
model.scheduler.set_timesteps(15, device=dev)
sig = model.scheduler.sigmas.contiguous()
g = torch.Generator(device=dev).manual_seed(1)
cond = (torch.randn(1, 16, T, H, W, device=dev, generator=g) * 0.5).bfloat16()
real_call = _lib.call

for B in [int(a) for a in sys.argv[1:]] or [1, 5]:
    ws = net._workspace(T, H, W, dev, None, batch=B)
    for b in range(B):
        net.prepare_condition(ws, cond, T, H, W, b)
        use_ca = net.prepare_context(ws, net.context_token(torch.full((1, 1), b % 5, dtype=torch.long, device=dev)), b)
    x = (torch.randn(B, 16, T, H, W, device=dev, generator=g).bfloat16() * sig[0]).bfloat16()
    for i in range(2):
        net.denoise_step(ws, x, sig[i:i + 1], sig[i + 1:i + 2], use_ca)
    torch.cuda.synchronize()
    # plain timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for i in range(n):
        net.denoise_step(ws, x, sig[2 + i:3 + i], sig[3 + i:4 + i], use_ca)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / n
    # per-call timing
    recs = []

    def timed_call(name, *args):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        real_call(name, *args)
        b_.record()
        key = name
        if name == "drb_gemm_bf16":
            key = f"gemm M={args[6]} N={args[7]} K={args[8]} epi={args[9]}"
        elif name.startswith("drb_gemm_qkv"):
            key = f"qkv_gemm M={args[6]}"
        recs.append((key, a, b_))

    _lib.call = timed_call
    net.denoise_step(ws, x, sig[6:7], sig[7:8], use_ca)
    torch.cuda.synchronize()
    _lib.call = real_call
    agg = collections.OrderedDict()
    for key, a, b_ in recs:
        t, c = agg.get(key, (0.0, 0))
        agg[key] = (t + a.elapsed_time(b_), c + 1)
    print(f"\n=== B = {B}: {total:.1f} ms per iteration = {total / B:.1f} ms per step; per-call sums (one iteration, events around every call):")
    for key, (t, c) in agg.items():
        print(f"  {key:44s} x{c:4d}  {t:9.2f} ms  = {t / B:8.2f} ms per step", flush=True)
    del ws, x
    net._ws.clear()
    torch.cuda.empty_cache()
