"""Environment-map conditioning of the forward renderer on B200 — drop-in for the two entry points of the reference
`preprocess_envmap.py` that `nodes.py` uses (`render_projection_from_panorama` :408-467, `tonemap_image_direct`
:469-526), with the same arguments, result dict and cache behaviour, but without nvdiffrast: the panorama -> 512x512
cube map -> projected lat-long view -> Reinhard / log / sRGB chain runs as three small sm_100a kernels
(csrc/envmap.cu).  HDR files on disk need `imageio` or OpenCV exactly as in the reference; ComfyUI IMAGE tensors do not.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Tuple, Union

import torch

from . import ops

CUBEMAP_RES = 512      # "official 512x512 cubemap" (reference :316)


class EnvironmentMapCache:
    """LRU cache keyed by (content hash, resolution, format, brightness, flip, rotation, device) (reference :23-66).  It
    holds the un-expanded (H,W,3) tone-mapped pair; the frame axis is added after the lookup, so the same environment map
    can serve clips of different length (the reference caches the expanded result and returns a stale frame count)."""

    def __init__(self, max_size: int = 10):
        self.cache: Dict[tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.max_size = max_size

    def get(self, key):
        value = self.cache.get(key)
        if value is not None:                       # most recently used goes last
            self.cache[key] = self.cache.pop(key)
        return value

    def put(self, key, value) -> None:
        if key in self.cache:
            self.cache.pop(key)
        elif len(self.cache) >= self.max_size:      # evict the least recently used entry, only for a NEW key
            self.cache.pop(next(iter(self.cache)))
        self.cache[key] = value

    def clear(self) -> None:
        self.cache.clear()


_env_cache = EnvironmentMapCache()


def compute_tensor_hash(tensor: torch.Tensor) -> str:
    t = tensor.detach()
    head = t.reshape(-1)[:4096].float().cpu().numpy().tobytes()
    return hashlib.md5(head + str(tuple(t.shape)).encode() + str(float(t.float().sum())).encode()).hexdigest()


def process_comfyui_tensor(tensor: torch.Tensor) -> torch.Tensor:
    """ComfyUI IMAGE (B,H,W,C) / (B,C,H,W) / (H,W,C) -> (H,W,3) (reference :247-261)"""
    if tensor.ndim == 4:
        if tensor.shape[1] in (3, 4):
            tensor = tensor.permute(0, 2, 3, 1)
        tensor = tensor[0]
    if tensor.shape[-1] == 4:
        tensor = tensor[..., :3]
    elif tensor.shape[-1] == 1:
        tensor = tensor.repeat(1, 1, 3)
    return tensor


def load_hdr_file(file_path: str) -> torch.Tensor:
    """.hdr / .exr -> float32 (H,W,3) (reference :208-245); needs imageio or OpenCV like the reference"""
    try:
        import imageio.v3 as iio
        img = iio.imread(file_path)
    except Exception:
        try:
            import cv2
            img = cv2.imread(file_path, cv2.IMREAD_UNCHANGED)
            if img is None:
                raise ValueError(file_path)
            img = img[..., ::-1].copy()
        except Exception as e:
            raise RuntimeError(f"cannot read HDR file {file_path}: neither imageio nor OpenCV could load it") from e
    t = torch.from_numpy(img).float()
    if t.ndim == 2:
        t = t.unsqueeze(-1).repeat(1, 1, 3)
    return t[..., :3]


def latlong_vec(res: Tuple[int, int], device="cuda") -> torch.Tensor:
    """unit direction of every lat-long pixel, (H,W,3) (reference :320-338); the env_nrm condition of the forward node"""
    import math
    H, W = res
    gy, gx = torch.meshgrid(torch.linspace(0.0 + 1.0 / H, 1.0 - 1.0 / H, H, device=device),
                            torch.linspace(-1.0 + 1.0 / W, 1.0 - 1.0 / W, W, device=device), indexing="ij")
    sintheta, costheta = torch.sin(gy * math.pi), torch.cos(gy * math.pi)
    sinphi, cosphi = torch.sin(gx * math.pi), torch.cos(gx * math.pi)
    return torch.stack((sintheta * sinphi, costheta, -sintheta * cosphi), dim=-1)


def _source(env_input: Union[str, torch.Tensor], device) -> torch.Tensor:
    if isinstance(env_input, str):
        img = load_hdr_file(env_input)
    elif isinstance(env_input, torch.Tensor):
        img = process_comfyui_tensor(env_input)
    else:
        raise ValueError(f"Unsupported input type: {type(env_input)}")
    return img.to(device=device, dtype=torch.float32).contiguous()


def _frames(ldr: torch.Tensor, lg: torch.Tensor, num_frames: int) -> Dict[str, torch.Tensor]:
    if num_frames > 1:
        return {"env_ldr": ldr.unsqueeze(0).expand(num_frames, -1, -1, -1), "env_log": lg.unsqueeze(0).expand(num_frames, -1, -1, -1)}
    return {"env_ldr": ldr.unsqueeze(0), "env_log": lg.unsqueeze(0)}


def _key(env_input, resolution, fmt, brightness, flip, rot, device):
    h = compute_tensor_hash(env_input) if isinstance(env_input, torch.Tensor) else hashlib.md5(str(env_input).encode()).hexdigest()
    return (h, tuple(resolution), fmt, float(brightness), bool(flip), float(rot), str(torch.device(device)))


def render_projection_from_panorama(env_input: Union[str, torch.Tensor], resolution: Tuple[int, int], env_brightness: float = 1.0,
                                    env_flip: bool = True, env_rot: float = 180.0, device="cuda", num_frames: int = 1,
                                    use_cache: bool = True, **kwargs) -> Dict[str, torch.Tensor]:
    """panorama -> cube map -> projected view -> {'env_ldr', 'env_log'}: (T,H,W,3) in [0,1] (reference :408-467)"""
    key = _key(env_input, resolution, "proj", env_brightness, env_flip, env_rot, device) if use_cache else None
    hit = _env_cache.get(key) if key is not None else None
    if hit is not None:
        return _frames(hit[0], hit[1], num_frames)
    H, W = resolution
    pano = _source(env_input, device)
    roll = int(pano.shape[1] * env_rot / 360) if env_rot != 0 else 0          # reference :282-284
    cube = ops.envmap_latlong_to_cubemap(pano, env_brightness, env_flip, roll, CUBEMAP_RES)
    ldr, lg = ops.envmap_project(cube, H, W)
    if key is not None:
        _env_cache.put(key, (ldr, lg))
    return _frames(ldr, lg, num_frames)


def tonemap_image_direct(env_input: Union[str, torch.Tensor], resolution: Tuple[int, int], device="cuda", num_frames: int = 1,
                         use_cache: bool = True, **kwargs) -> Dict[str, torch.Tensor]:
    """pre-rendered HDR probe image -> resize + tone mapping (reference :469-526)"""
    key = _key(env_input, resolution, "ball", 1.0, False, 0.0, device) if use_cache else None
    hit = _env_cache.get(key) if key is not None else None
    if hit is not None:
        return _frames(hit[0], hit[1], num_frames)
    H, W = resolution
    ldr, lg = ops.envmap_tonemap(_source(env_input, device), H, W)
    if key is not None:
        _env_cache.put(key, (ldr, lg))
    return _frames(ldr, lg, num_frames)


def clear_environment_cache() -> None:
    _env_cache.clear()


def get_cache_stats() -> Dict[str, int]:
    return {"cache_size": len(_env_cache.cache), "max_size": _env_cache.max_size}
