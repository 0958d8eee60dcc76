"""ComfyUI node classes of DiffusionRenderer on B200 — the reference's registry (`nodes.py:335-347`) with the same
class names, INPUT_TYPES / RETURN_TYPES / FUNCTION / CATEGORY and method signatures, driving the B200 pipeline.

ComfyUI (`comfy`, `folder_paths`) is imported lazily inside the methods that need it, so the classes import (and are
testable) without a ComfyUI host.  Environment maps are projected / tone-mapped on the device by this package's
`preprocess_envmap.py` (no nvdiffrast); `Cosmos1ForwardRenderer` also accepts pre-computed `env_ldr` / `env_log` tensors
in a dict.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from .diffusion_renderer_config import get_inverse_renderer_config
from .diffusion_renderer_pipeline import CleanDiffusionRendererPipeline
from .model_diffusion_renderer import CleanDiffusionRendererModel

# cosmos_predict1 .../rendering_utils.py mapping, reference nodes.py:35-41
GBUFFER_INDEX_MAPPING = {"basecolor": 0, "metallic": 1, "roughness": 2, "normal": 3, "depth": 4}
INFERENCE_PASSES = ("basecolor", "metallic", "roughness", "normal", "depth")


def _to_5d(t, name="image"):
    """IMAGE -> (B,T,H,W,C): 3-D adds batch and time, 4-D (B,H,W,C) adds time, list is stacked (reference :156-177)."""
    if isinstance(t, list):
        try:
            return torch.stack(t, dim=0)
        except Exception:
            return t[0].unsqueeze(0)
    if isinstance(t, torch.Tensor):
        if t.ndim == 3:
            return t.unsqueeze(0).unsqueeze(0)
        if t.ndim == 4:
            return t.unsqueeze(1)
        if t.ndim == 5:
            return t
        raise ValueError(f"Unsupported tensor dimension for '{name}': {t.ndim}. Expected 3D, 4D, or 5D.")
    raise TypeError(f"Unsupported input type for '{name}': {type(t)}. Expected torch.Tensor or list of Tensors.")


def latlong_vec(res, device=None):
    """Unit direction of every lat-long pixel, (H,W,3) (reference preprocess_envmap.py:320-338)."""
    gy, gx = torch.meshgrid(torch.linspace(0.0 + 1.0 / res[0], 1.0 - 1.0 / res[0], res[0], device=device),
                            torch.linspace(-1.0 + 1.0 / res[1], 1.0 - 1.0 / res[1], res[1], device=device), indexing="ij")
    sintheta, costheta = torch.sin(gy * math.pi), torch.cos(gy * math.pi)
    sinphi, cosphi = torch.sin(gx * math.pi), torch.cos(gx * math.pi)
    return torch.stack((sintheta * sinphi, costheta, -sintheta * cosphi), dim=-1)


def _net_dims_from_state_dict(sd, defaults) -> dict:
    """model_channels / num_blocks / num_heads of the GeneralDIT a checkpoint holds (head_dim is 128)"""
    w = sd.get("net.affline_norm.weight")
    if w is None:
        return {}
    D = int(w.shape[0])
    blocks = {k.split(".")[2] for k in sd if k.startswith("net.blocks.block")}
    return {"model_channels": D, "num_blocks": len(blocks) or defaults["num_blocks"], "num_heads": D // 128}


class LoadDiffusionRendererModel:
    @classmethod
    def INPUT_TYPES(s):
        import folder_paths
        return {"required": {"model": (folder_paths.get_filename_list("diffusion_models"),
                                       {"tooltip": "Models are loaded from 'ComfyUI/models/diffusion_models'"})}}

    RETURN_TYPES = ("DIFFUSION_RENDERER_PIPELINE",)
    FUNCTION = "load_pipeline"
    CATEGORY = "Cosmos1"

    def load_pipeline(self, model):
        import comfy.model_management as mm
        import comfy.utils
        import folder_paths
        from .CleanVAE import CleanVAE
        device, dtype = mm.get_torch_device(), torch.bfloat16
        vae_dir = os.path.join(folder_paths.models_dir, "vae", "Cosmos-1.0-Tokenizer-CV8x8x8", "vae")
        if not os.path.isdir(vae_dir):
            raise FileNotFoundError(f"Image VAE subfolder not found at: {vae_dir}")
        vae = CleanVAE(model_path=vae_dir)
        vae.to(device)
        vae.reset_dtype(dtype)
        ckpt = folder_paths.get_full_path("diffusion_models", model)
        sd = comfy.utils.load_torch_file(ckpt, safe_load=True)
        if "model" in sd:
            sd = sd["model"]
        # the checkpoint tells which renderer it is: only the inverse net has a context embedding — and how
        # large the net is (the released checkpoints are the 7B FADITV2 net of the reference config, nodes.py:94-96)
        from .diffusion_renderer_config import get_forward_renderer_config
        cfg = get_inverse_renderer_config() if "net.context_embedding.weight" in sd else get_forward_renderer_config()
        cfg["net"].update(_net_dims_from_state_dict(sd, cfg["net"]))
        with torch.device("meta"):
            m = CleanDiffusionRendererModel(cfg)
        m.to_empty(device=device)
        m.to(dtype=dtype)
        m.load_state_dict(sd, strict=True)
        del sd
        mm.soft_empty_cache()
        m.eval()
        return (CleanDiffusionRendererPipeline(checkpoint_dir=os.path.dirname(ckpt), checkpoint_name=os.path.basename(ckpt),
                                               model_type=None, vae_instance=vae, model_instance=m, guidance=0.0,
                                               num_steps=15, seed=42),)


class Cosmos1InverseRenderer:
    @classmethod
    def INPUT_TYPES(s):
        return {"required": {"pipeline": ("DIFFUSION_RENDERER_PIPELINE",), "image": ("IMAGE",)},
                "optional": {"guidance": ("FLOAT", {"default": 0.0, "min": 0.0, "max": 10.0, "step": 0.1}),
                             "seed": ("INT", {"default": 42, "min": 0, "max": 0xffffffffffffffff})}}

    RETURN_TYPES = ("IMAGE", "IMAGE", "IMAGE", "IMAGE", "IMAGE")
    RETURN_NAMES = ("base_color", "metallic", "roughness", "normal", "depth")
    FUNCTION = "run_inverse_pass"
    CATEGORY = "Cosmos1"

    def run_inverse_pass(self, pipeline, image, guidance=0.0, seed=42):
        pipeline.set_model_type("inverse")
        pipeline.guidance = guidance
        pipeline.seed = seed
        pipeline.pinned_output = True          # the frames are converted to float tensors below, before the next call
        clip = _to_5d(image).permute(0, 4, 1, 2, 3) * 2.0 - 1.0          # (B,3,T,H,W) in [-1,1]
        try:
            from comfy.utils import ProgressBar
            pbar = ProgressBar(len(INFERENCE_PASSES))
        except Exception:
            pbar = None
        outputs = {}

        def keep(name, arr):
            if arr is None:            # a helper rank of a context-parallel group: the frames went to the output rank
                outputs[name] = None
                return
            out = torch.from_numpy(arr).float() / 255.0
            b, t, h, w, c = out.shape
            outputs[name] = out.reshape(b * t, h, w, c)
            if pbar is not None:
                pbar.update(1)

        if getattr(pipeline, "wants_batched_passes", lambda: False)() and clip.shape[0] == 1:
            # one sampler run for all five passes (batched along the token axis); same values as the loop below
            arrs = pipeline.generate_video_passes({"rgb": clip, "video": clip}, [GBUFFER_INDEX_MAPPING[n] for n in INFERENCE_PASSES],
                                                  normalize_normal=[n == "normal" for n in INFERENCE_PASSES], seed=seed)
            for name, arr in zip(INFERENCE_PASSES, arrs):
                keep(name, arr)
            return tuple(outputs[n] for n in INFERENCE_PASSES)
        with pipeline.shared_conditions():                               # the clip is tokenised once for all five passes
            for name in INFERENCE_PASSES:
                batch = {"rgb": clip, "video": clip,
                         "context_index": torch.full((clip.shape[0], 1), GBUFFER_INDEX_MAPPING[name], dtype=torch.long)}
                keep(name, pipeline.generate_video(data_batch=batch, normalize_normal=(name == "normal"), seed=seed))
        return tuple(outputs[n] for n in INFERENCE_PASSES)


class Cosmos1ForwardRenderer:
    @classmethod
    def INPUT_TYPES(s):
        img = ("IMAGE",)
        return {"required": {"pipeline": ("DIFFUSION_RENDERER_PIPELINE",), "depth": img, "normal": img, "roughness": img,
                             "metallic": img, "base_color": img, "env_map": img},
                "optional": {"guidance": ("FLOAT", {"default": 0.0, "min": 0.0, "max": 2.0, "step": 0.1}),
                             "seed": ("INT", {"default": 42, "min": 0, "max": 0xffffffffffffffff}),
                             "env_format": (["proj", "ball"], {"default": "proj"}),
                             "env_brightness": ("FLOAT", {"default": 1.0, "min": 0.0, "max": 2.0, "step": 0.1}),
                             "env_flip_horizontal": ("BOOLEAN", {"default": False}),
                             "env_rotation": ("FLOAT", {"default": 180.0, "min": 0, "max": 360, "step": 1.0})}}

    RETURN_TYPES = ("IMAGE",)
    FUNCTION = "run_forward_pass"
    CATEGORY = "Cosmos1"

    @staticmethod
    def _environment(env_map, H, W, T, device, env_format, env_brightness, env_flip, env_rot):
        """{'env_ldr','env_log'}: (T,H,W,3) in [0,1] (reference nodes.py:283-298).  A dict input is taken as already
        projected / tone-mapped; otherwise the B200 preprocess_envmap module does it on the device (no nvdiffrast)."""
        if isinstance(env_map, dict) and "env_ldr" in env_map and "env_log" in env_map:
            return env_map
        from . import preprocess_envmap as pe
        if env_format == "proj":
            return pe.render_projection_from_panorama(env_input=env_map, resolution=(H, W), num_frames=T, device=device,
                                                      env_brightness=env_brightness, env_flip=env_flip, env_rot=env_rot)
        return pe.tonemap_image_direct(env_input=env_map, resolution=(H, W), num_frames=T, device=device)

    def run_forward_pass(self, pipeline, depth, normal, roughness, metallic, base_color, env_map, guidance=0.0, seed=42,
                         env_format="proj", env_brightness=1.0, env_flip_horizontal=False, env_rotation=0.0):
        pipeline.set_model_type("forward")
        pipeline.guidance = guidance
        pipeline.seed = seed
        pipeline.pinned_output = True          # the frames are converted to a float tensor below, before the next call
        g5 = {n: _to_5d(t, n) for n, t in (("depth", depth), ("normal", normal), ("roughness", roughness),
                                            ("metallic", metallic), ("base_color", base_color))}
        B, T, H, W, _ = g5["depth"].shape
        keymap = {"base_color": "basecolor", "depth": "depth", "normal": "normal", "roughness": "roughness", "metallic": "metallic"}
        batch = {keymap[n]: t.permute(0, 4, 1, 2, 3) * 2.0 - 1.0 for n, t in g5.items()}
        batch["video"] = batch["depth"]
        dev = g5["depth"].device
        env = self._environment(env_map, H, W, T, getattr(pipeline, "device", torch.device("cuda")), env_format, env_brightness,
                                env_flip_horizontal, env_rotation)
        batch["env_ldr"] = (env["env_ldr"].permute(3, 0, 1, 2).unsqueeze(0) * 2.0 - 1.0).expand(B, -1, -1, -1, -1)
        batch["env_log"] = (env["env_log"].permute(3, 0, 1, 2).unsqueeze(0) * 2.0 - 1.0).expand(B, -1, -1, -1, -1)
        # the reference passes `resolution=` to a parameter named `res` (nodes.py:300, defect D2): positional here
        nrm = latlong_vec((H, W), device=dev).permute(2, 0, 1).unsqueeze(0).unsqueeze(2)
        batch["env_nrm"] = nrm.expand(B, -1, T, -1, -1)
        arr = pipeline.generate_video(data_batch=batch, seed=seed)
        return (torch.from_numpy(arr).float() / 255.0,)


class LoadHDRImage:
    @classmethod
    def INPUT_TYPES(s):
        return {"required": {"path": ("STRING", {"tooltip": "Path to HDR image (.hdr, .exr)"})}}

    RETURN_TYPES = ("IMAGE",)
    FUNCTION = "load_hdr"
    CATEGORY = "Cosmos1"

    def load_hdr(self, path):
        import imageio
        img = imageio.imread(path, format="HDR-FI")
        if img.ndim == 2:
            img = np.stack([img] * 3, axis=-1)
        elif img.ndim == 3 and img.shape[2] == 1:
            img = np.repeat(img, 3, axis=2)
        return (torch.from_numpy(img).float().unsqueeze(0),)


NODE_CLASS_MAPPINGS = {
    "LoadDiffusionRendererModel": LoadDiffusionRendererModel,
    "Cosmos1InverseRenderer": Cosmos1InverseRenderer,
    "Cosmos1ForwardRenderer": Cosmos1ForwardRenderer,
    "LoadHDRImage": LoadHDRImage,
}
NODE_DISPLAY_NAME_MAPPINGS = {
    "LoadDiffusionRendererModel": "Load Diffusion Renderer Model",
    "Cosmos1InverseRenderer": "Cosmos1 Inverse Renderer",
    "Cosmos1ForwardRenderer": "Cosmos1 Forward Renderer",
    "LoadHDRImage": "Load HDR Image",
}
