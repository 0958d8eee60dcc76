"""Generate tests/golden/*.npz by running the REAL reference (from /root/reference) on CPU, fp32.

Run in the build container only:  python tests/golden/make_golden.py
The reference has no golden vectors of its own; these are produced by importing
its modules through oracle/ref_loader.py (one documented patch, defect D1) and are what the
`-m "not gpu"` tests check the oracle restatement against on machines without /root/reference.
Weights are NOT stored: they are regenerated from oracle.weights.make_state_dict(dims, seed).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader                                            # noqa: E402
from oracle.vae_stub import StubVAE                                      # noqa: E402
from oracle.weights import MICRO_FORWARD, MICRO_INVERSE, make_state_dict, net_only  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHT_SEED = 3
LAT = (2, 8, 12)          # latent T,H,W  -> patches 2x4x6 = 48 tokens
PIX = (9, 64, 96)         # pixel  T,H,W


def dit_case(dit, dims, tag):
    sd = make_state_dict(dims, seed=WEIGHT_SEED)
    net = dit.CleanDiffusionRendererGeneralDIT(**dims.net_kwargs())
    net.load_state_dict(net_only(sd), strict=True)
    net.eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 16, *LAT, generator=g)
    cond = torch.randn(1, dims.additional_concat_ch, *LAT, generator=g)
    out = {}
    for k, (sigma, ci) in enumerate(((80.0, 0), (1.2589254, 3), (0.02, 4))):
        with torch.no_grad():
            y = net(x=x, timesteps=torch.tensor(sigma), latent_condition=cond,
                    context_index=torch.full((1, 1), ci, dtype=torch.long))
        out[f"{tag}_sigma{k}"] = np.float32(sigma)
        out[f"{tag}_ci{k}"] = np.int64(ci)
        out[f"{tag}_F{k}"] = y.numpy()
    out[f"{tag}_x"] = x.numpy()
    out[f"{tag}_cond"] = cond.numpy()
    return out


def sampler_case(cfgm, mdl, pipem, dims, tag, model_type):
    """Full reference generate_video() on CPU fp32 with the stub tokenizer: 2 Euler steps."""
    T, H, W = PIX
    getc = cfgm.get_inverse_renderer_config if model_type == "inverse" else cfgm.get_forward_renderer_config
    config = getc(H, W, T)
    config["net"].update(model_channels=dims.model_channels, num_blocks=dims.num_blocks, num_heads=dims.num_heads)
    config["model_type"] = model_type
    model = mdl.CleanDiffusionRendererModel(config)
    sd = make_state_dict(dims, seed=WEIGHT_SEED)
    model.load_state_dict(sd, strict=True)
    P = pipem.CleanDiffusionRendererPipeline("", "none", model_type=None, vae_instance=StubVAE(),
                                             model_instance=model, guidance=0.0, num_steps=2, seed=42,
                                             dtype=torch.float32)
    P.device = torch.device("cpu")
    P.set_model_type(model_type)
    g = torch.Generator().manual_seed(1234)     # inputs are NOT stored: tests redraw them from this seed
    out = {}
    if model_type == "inverse":
        clip = torch.rand(1, 3, T, H, W, generator=g) * 2 - 1
        batch = {"rgb": clip, "video": clip, "context_index": torch.full((1, 1), 3, dtype=torch.long)}
    else:
        keys = ["basecolor", "normal", "metallic", "roughness", "depth", "env_ldr", "env_log", "env_nrm"]
        batch = {k: torch.rand(1, 3, T, H, W, generator=g) * 2 - 1 for k in keys}
        batch["video"] = batch["depth"]
        batch["context_index"] = torch.zeros((1, 1), dtype=torch.long)     # defect D4: required positionally
    # latent after sampling (before decode) and final frames
    state_shape = [16, (T - 1) // 8 + 1, H // 8, W // 8]
    P._ensure_model_loaded(tuple(batch["video"].shape))
    latent = P.model.generate_samples_from_batch(dict(batch), guidance=0.0, state_shape=state_shape,
                                                 num_steps=2, seed=42)
    frames = P.generate_video(dict(batch), normalize_normal=(model_type == "inverse"), seed=42)
    out[f"{tag}_latent"] = latent.numpy()
    out[f"{tag}_frames"] = frames
    # guidance > 0 path (CFG, model_diffusion_renderer.py:230-232)
    latent_g = P.model.generate_samples_from_batch(dict(batch), guidance=2.0, state_shape=state_shape,
                                                   num_steps=2, seed=42)
    out[f"{tag}_latent_cfg2"] = latent_g.numpy()
    return out


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)                       # fixed reduction order
    torch.use_deterministic_algorithms(True)
    dit, cfgm, mdl, pipem = ref_loader.load()
    out = {}
    out.update(dit_case(dit, MICRO_INVERSE, "inv"))
    out.update(dit_case(dit, MICRO_FORWARD, "fwd"))
    np.savez_compressed(os.path.join(HERE, "dit_micro.npz"), **out)
    out = {}
    out.update(sampler_case(cfgm, mdl, pipem, MICRO_INVERSE, "inv", "inverse"))
    out.update(sampler_case(cfgm, mdl, pipem, MICRO_FORWARD, "fwd", "forward"))
    np.savez_compressed(os.path.join(HERE, "sampler_micro.npz"), **out)
    for f in ("dit_micro.npz", "sampler_micro.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
