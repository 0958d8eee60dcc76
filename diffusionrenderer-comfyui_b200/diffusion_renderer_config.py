"""Plain-dict configuration of the two DiffusionRenderer models — same function names, keys and values as the
reference `diffusion_renderer_config.py` (get_network_config:47, get_scheduler_config:106, get_vae_config:121,
get_inverse_renderer_config:131, get_forward_renderer_config:191, get_config_from_tensor_shape:277,
validate_config:308, PRESET_CONFIGS:352), rebuilt around one table of per-renderer differences.  Host logic only.
"""
from __future__ import annotations

import copy
from typing import Any, Dict

_NETWORK = {
    # FADITV2_7B
    "model_channels": 4096, "num_blocks": 28, "num_heads": 32, "head_dim": 128, "mlp_ratio": 4.0, "context_dim": 1024,
    "adaln_lora_dim": 256, "time_embed_dim": 4096, "max_time_embed_period": 10000,
    "in_channels": 16, "out_channels": 16, "patch_spatial": 2, "patch_temporal": 1,
    "max_img_h": 240, "max_img_w": 240, "max_frames": 128,
    "block_config": "FA-CA-MLP", "concat_padding_mask": True, "block_x_format": "THWBD",
    "pos_emb_cls": "rope3d", "pos_emb_learnable": False, "pos_emb_interpolation": "crop",
    "rope_h_extrapolation_ratio": 1.0, "rope_w_extrapolation_ratio": 1.0, "rope_t_extrapolation_ratio": 2.0,
    "affline_emb_norm": True, "use_adaln_lora": True, "extra_per_block_abs_pos_emb": True,
    "extra_per_block_abs_pos_emb_type": "sincos", "extra_h_extrapolation_ratio": 1.0, "extra_w_extrapolation_ratio": 1.0,
    "extra_t_extrapolation_ratio": 1.0,
    "crossattn_emb_channels": 1024,   # the context embedding of the released checkpoints is 1024-wide
}

_SCHEDULER = {
    "type": "EDMEulerScheduler", "sigma_max": 80.0, "sigma_min": 0.02, "sigma_data": 0.5, "num_train_timesteps": 1000,
    "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear", "prediction_type": "v_prediction",
}

# what differs between the two renderers (reference :158-170 and :218-234)
_RENDERERS = {
    "inverse": {
        "condition_keys": ["rgb"], "condition_drop_rate": 0.1, "append_condition_mask": False,
        "net": {"additional_concat_ch": 16, "use_context_embedding": True},
    },
    "forward": {
        "condition_keys": ["basecolor", "normal", "metallic", "roughness", "depth", "env_ldr", "env_log", "env_nrm"],
        "condition_drop_rate": 0.05, "append_condition_mask": True,
        "net": {"additional_concat_ch": 17 * 8, "use_context_embedding": False},   # 8 x (16 latent + 1 mask) = 136
    },
}


class CleanDiffusionRendererConfig:
    """Base values shared by both renderers (reference :24-44)."""

    def __init__(self):
        self.sigma_data = 0.5
        self.precision = "bfloat16"
        self.input_data_key = "video"
        self.latent_shape = [16, 8, 88, 160]
        self.condition_keys = ["rgb"]
        self.condition_drop_rate = 0.0
        self.append_condition_mask = True
        self.model_channels = 4096
        self.num_blocks = 28
        self.num_heads = 32


def get_network_config() -> Dict[str, Any]:
    return dict(_NETWORK)


def get_scheduler_config() -> Dict[str, Any]:
    return dict(_SCHEDULER)


def get_vae_config(num_frames: int = 57) -> Dict[str, Any]:
    return {"pixel_chunk_duration": num_frames, "latent_channels": 16, "spatial_compression_ratio": 8,
            "temporal_compression_ratio": 8}


def _renderer_config(kind: str, height: int, width: int, num_frames: int) -> Dict[str, Any]:
    base, spec = CleanDiffusionRendererConfig(), _RENDERERS[kind]
    return {
        "sigma_data": base.sigma_data, "precision": base.precision, "input_data_key": base.input_data_key,
        # informational only, and `num_frames // 8 + 1` disagrees with the pipeline's (T-1)//8+1 when T % 8 == 0
        # (reference :147 vs pipeline :275); kept as the reference computes it
        "latent_shape": [16, num_frames // 8 + 1, height // 8, width // 8],
        "condition_keys": list(spec["condition_keys"]), "condition_drop_rate": spec["condition_drop_rate"],
        "append_condition_mask": spec["append_condition_mask"],
        "net": {**get_network_config(), **spec["net"], "crossattn_emb_channels": 1024},
        "scheduler": get_scheduler_config(), "vae": get_vae_config(num_frames),
        "guidance": 2.0, "num_steps": 20, "height": height, "width": width, "num_video_frames": num_frames,
    }


def get_inverse_renderer_config(height: int = 704, width: int = 1280, num_frames: int = 57) -> Dict[str, Any]:
    """RGB video -> one G-buffer per pass (selected by context_index)."""
    return _renderer_config("inverse", height, width, num_frames)


def get_forward_renderer_config(height: int = 704, width: int = 1280, num_frames: int = 57) -> Dict[str, Any]:
    """G-buffers + environment map -> RGB video."""
    return _renderer_config("forward", height, width, num_frames)


def get_config_by_model_type(model_type: str, height: int = 704, width: int = 1280, num_frames: int = 57) -> Dict[str, Any]:
    kind = model_type.lower()
    if kind not in _RENDERERS:
        raise ValueError(f"Unknown model type: {model_type}. Must be 'inverse' or 'forward'")
    return _renderer_config(kind, height, width, num_frames)


def get_config_from_tensor_shape(model_type, tensor_shape):
    if len(tensor_shape) != 5:
        raise ValueError(f"Expected a 5D tensor shape, but got {len(tensor_shape)} dimensions.")
    _, _, T, H, W = tensor_shape
    if model_type not in _RENDERERS:
        raise ValueError(f"Unknown model type for config generation: {model_type}")
    return _renderer_config(model_type, H, W, T)


def validate_config(config: Dict[str, Any]) -> None:
    for key in ("sigma_data", "precision", "input_data_key", "latent_shape", "condition_keys", "net", "scheduler", "vae"):
        if key not in config:
            raise ValueError(f"Missing required config key: {key}")
    latent_shape = config["latent_shape"]
    if not isinstance(latent_shape, list) or len(latent_shape) != 4:
        raise ValueError(f"Invalid latent_shape: {latent_shape}. Expected [C, T, H, W] format.")
    for key in ("model_channels", "num_blocks", "num_heads", "in_channels", "out_channels"):
        if key not in config["net"]:
            raise ValueError(f"Missing required net config key: {key}")


PRESET_CONFIGS = {
    "inverse_1024x1024": get_inverse_renderer_config(1024, 1024, 1),
    "forward_1024x1024": get_forward_renderer_config(1024, 1024, 1),
    "inverse_704x1280_video": get_inverse_renderer_config(704, 1280, 57),
    "forward_704x1280_video": get_forward_renderer_config(704, 1280, 57),
}


def get_preset_config(preset_name: str) -> Dict[str, Any]:
    if preset_name not in PRESET_CONFIGS:
        raise ValueError(f"Unknown preset: {preset_name}. Available: {list(PRESET_CONFIGS.keys())}")
    return copy.deepcopy(PRESET_CONFIGS[preset_name])
