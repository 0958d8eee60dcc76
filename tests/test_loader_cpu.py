"""Checkpoint loading through the ComfyUI loader node (reference nodes.py:74-127) with stubbed `comfy` / `folder_paths`:
a `.safetensors` state dict and a torch checkpoint wrapped as {"model": sd} (:98-101), plus a tokenizer directory in the
diffusers layout (config.json + diffusion_pytorch_model.safetensors, :83-92) — meta skeleton -> to_empty -> strict load
(:103-115).  Runs on the CPU: loading involves no kernel."""
import json
import os
import sys
import types

import pytest
import torch

from oracle import vae_oracle as vo
from oracle.weights import MICRO_FORWARD, MICRO_INVERSE, make_state_dict


def _stub_comfy(monkeypatch, models_dir, device):
    folder_paths = types.ModuleType("folder_paths")
    folder_paths.models_dir = str(models_dir)
    folder_paths.get_filename_list = lambda kind: sorted(os.listdir(os.path.join(models_dir, kind)))
    folder_paths.get_full_path = lambda kind, name: os.path.join(models_dir, kind, name)
    comfy = types.ModuleType("comfy")
    mm = types.ModuleType("comfy.model_management")
    mm.get_torch_device = lambda: torch.device(device)
    mm.soft_empty_cache = lambda: None
    utils = types.ModuleType("comfy.utils")

    def load_torch_file(path, safe_load=False):
        if path.endswith(".safetensors"):
            from safetensors.torch import load_file
            return load_file(path)
        return torch.load(path, map_location="cpu", weights_only=safe_load)
    utils.load_torch_file = load_torch_file
    comfy.model_management, comfy.utils = mm, utils
    for name, mod in (("folder_paths", folder_paths), ("comfy", comfy), ("comfy.model_management", mm), ("comfy.utils", utils)):
        monkeypatch.setitem(sys.modules, name, mod)


def write_models(root, dims, wrapped: bool, fname: str, vae_dims=vo.SMALL_VAE):
    """a ComfyUI models/ tree: diffusion_models/<fname> and vae/Cosmos-1.0-Tokenizer-CV8x8x8/vae/ (diffusers layout)"""
    from safetensors.torch import save_file
    dm = os.path.join(root, "diffusion_models")
    vd = os.path.join(root, "vae", "Cosmos-1.0-Tokenizer-CV8x8x8", "vae")
    os.makedirs(dm, exist_ok=True)
    os.makedirs(vd, exist_ok=True)
    sd = {k: v.contiguous() for k, v in make_state_dict(dims, seed=5, dtype=torch.bfloat16).items()}
    if fname.endswith(".safetensors"):
        save_file(sd, os.path.join(dm, fname))
    else:
        torch.save({"model": sd} if wrapped else sd, os.path.join(dm, fname))
    vsd = vo.make_vae_state_dict(vae_dims, seed=7)
    save_file({k: v.contiguous() for k, v in vsd.items()}, os.path.join(vd, "diffusion_pytorch_model.safetensors"))
    with open(os.path.join(vd, "config.json"), "w") as f:
        json.dump({"_class_name": "AutoencoderKLCosmos", "encoder_block_out_channels": list(vae_dims.encoder_block_out_channels),
                   "decode_block_out_channels": list(vae_dims.decode_block_out_channels), "latent_channels": 16, "patch_size": 4,
                   "patch_type": "haar", "latents_mean": [0.0] * 256, "latents_std": [1.0] * 256}, f)
    return sd, vsd


@pytest.mark.parametrize("dims,fname,wrapped", [(MICRO_INVERSE, "inverse.safetensors", False), (MICRO_FORWARD, "forward.pt", True),
                                                (MICRO_INVERSE, "inverse_flat.pt", False)])
def test_loader_node_reads_checkpoint_files(tmp_path, monkeypatch, dims, fname, wrapped):
    from drb200 import nodes
    sd, vsd = write_models(str(tmp_path), dims, wrapped, fname)
    _stub_comfy(monkeypatch, str(tmp_path), "cpu")
    assert fname in nodes.LoadDiffusionRendererModel.INPUT_TYPES()["required"]["model"][0]
    (pipe,) = nodes.LoadDiffusionRendererModel().load_pipeline(fname)
    model = pipe.pre_loaded_model_instance
    assert pipe.guidance == 0.0 and pipe.num_steps == 15 and pipe.seed == 42 and pipe.model_type is None      # nodes.py:117-127
    assert model.net.model_channels == dims.model_channels and model.net.num_blocks == dims.num_blocks
    assert hasattr(model.net, "context_embedding") == dims.use_context_embedding
    got = model.state_dict()
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert got[k].dtype == torch.bfloat16 and torch.equal(got[k], v), k
    vgot = pipe.vae_instance.model.state_dict()
    assert set(vgot) == set(vsd)
    for k, v in vsd.items():
        assert torch.equal(vgot[k].float(), v.bfloat16().float()), k
    assert pipe.vae_instance.config.latents_mean == [0.0] * 256


def test_loader_node_errors(tmp_path, monkeypatch):
    from drb200 import nodes
    os.makedirs(tmp_path / "diffusion_models")
    _stub_comfy(monkeypatch, str(tmp_path), "cpu")
    with pytest.raises(FileNotFoundError):                       # no tokenizer directory (nodes.py:85-86)
        nodes.LoadDiffusionRendererModel().load_pipeline("x.safetensors")
    sd, _ = write_models(str(tmp_path), MICRO_INVERSE, False, "broken.pt")
    bad = dict(sd)
    bad.pop("net.final_layer.linear.weight")
    torch.save(bad, tmp_path / "diffusion_models" / "broken.pt")
    with pytest.raises(RuntimeError):                            # strict load (nodes.py:107)
        nodes.LoadDiffusionRendererModel().load_pipeline("broken.pt")
