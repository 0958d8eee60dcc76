"""Two full 7B denoise steps (one warm-up, one measured) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import diffusion_renderer_config as cfgm
from drb200.model_diffusion_renderer import CleanDiffusionRendererModel

dev = torch.device("cuda", 0)
cfg = cfgm.get_inverse_renderer_config(704, 1280, 57)
cfg["model_type"] = "inverse"
with torch.device("meta"):
    model = CleanDiffusionRendererModel(cfg)
model = model.to_empty(device=dev).to(torch.bfloat16)
net = model.net.init_weights_(seed=0)
t, h, w = 8, 88, 160
g = torch.Generator(device=dev).manual_seed(1234)
cond = (torch.randn(1, 16, t, h, w, device=dev, generator=g) * 0.5).bfloat16()
model.scheduler.set_timesteps(15, device=dev)
sig = model.scheduler.sigmas.contiguous()
x = (torch.randn(16, t, h, w, device=dev, generator=g).bfloat16() * sig[0]).bfloat16()
net._ensure_packed()
ws = net._workspace(t, h, w, dev)
net.prepare_condition(ws, cond, t, h, w)
use_ca = net.prepare_context(ws, net.context_token(torch.zeros(1, 1, dtype=torch.long, device=dev)))
for i in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    net.denoise_step(ws, x, sig[i:i + 1], sig[i + 1:i + 2], use_ca)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {i}: {e0.elapsed_time(e1):.1f} ms")
