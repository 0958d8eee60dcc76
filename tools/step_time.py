"""Sustained ms/step of the 7B denoise step in this process (same-box comparisons of environment-selected kernel variants:
run it several times in ONE gpurun call).  usage: python tools/step_time.py [rounds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from drb200 import diffusion_renderer_config as cfgm
from drb200.model_diffusion_renderer import CleanDiffusionRendererModel

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
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
x0 = (torch.randn(16, t, h, w, device=dev, generator=g).bfloat16() * sig[0]).bfloat16()
net._ensure_packed()
ws = net._workspace(t, h, w, dev)
net.prepare_condition(ws, cond, t, h, w)
use_ca = net.prepare_context(ws, net.context_token(torch.zeros(1, 1, dtype=torch.long, device=dev)))


def run(n):
    x = x0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        net.denoise_step(ws, x, sig[i:i + 1], sig[i + 1:i + 2], use_ca)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


run(6)
res = [run(8) for _ in range(rounds)]
print(f"DRB_ATTN_POLY={os.environ.get('DRB_ATTN_POLY', 'default')}: " + " ".join(f"{r:.1f}" for r in res) + " ms/step", flush=True)
