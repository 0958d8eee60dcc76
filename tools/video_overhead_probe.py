"""Where the non-DiT time of generate_video goes (one G-buffer pass, one Euler step): H2D + cast of the fp32 clip, tokenizer
encode, sampler set-up, decode + post-process, D2H of the uint8 frames.  Small 7B-free model: the DiT is tiny here on purpose.

    python tools/video_overhead_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from drb200 import diffusion_renderer_config as cfgm
from drb200.diffusion_renderer_pipeline import CleanDiffusionRendererPipeline
from drb200.model_diffusion_renderer import CleanDiffusionRendererModel

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
f, hh, ww = 57, 704, 1280
cfg = cfgm.get_inverse_renderer_config(hh, ww, f)
cfg["model_type"] = "inverse"
cfg["net"].update(model_channels=512, num_blocks=1, num_heads=4)
with torch.device("meta"):
    model = CleanDiffusionRendererModel(cfg)
model = model.to_empty(device=dev).to(torch.bfloat16)
model.net.init_weights_(seed=0)
vae = bench.random_tokenizer(torch, dev)
pipe = CleanDiffusionRendererPipeline(checkpoint_dir="", checkpoint_name="", model_type="inverse", vae_instance=vae, model_instance=model,
                                      guidance=0.0, num_steps=1, seed=42)
clip = (torch.rand(1, 3, f, hh, ww, generator=torch.Generator().manual_seed(1234)) * 2 - 1).pin_memory()


def timed(label, fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    print(f"{label:58s} {(time.perf_counter() - t0) / n * 1e3:8.1f} ms", flush=True)
    return out


batch = {"rgb": clip, "video": clip, "context_index": torch.zeros(1, 1, dtype=torch.long)}
timed("generate_video, 1 pass, 1 step (tiny DiT)", lambda: pipe.generate_video(batch, seed=42))
d = timed("H2D of the pinned fp32 clip + cast to bf16 (x2 keys)", lambda: pipe._move_to_device(batch))
z = timed("tokenizer encode (scaled)", lambda: model.encode(d["rgb"]))
fr = timed("decode + post-process (fused uint8 store)", lambda: pipe._decode_frames(model, z, False))
timed("frames.cpu() (pageable destination)", lambda: fr.cpu())
pin = torch.empty(fr.shape, dtype=torch.uint8).pin_memory()
timed("frames -> pinned host buffer", lambda: pin.copy_(fr, non_blocking=True))
timed("frames.cpu().numpy()", lambda: fr.cpu().numpy())
