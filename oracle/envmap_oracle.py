"""Torch restatement of the reference environment-map preprocessing (preprocess_envmap.py).

TEST INFRASTRUCTURE (see oracle/__init__.py) — never imported by the product.

The pure-torch stages follow the reference line by line and are pinned against the reference's own functions in
tests/test_oracle_vs_reference.py (the module is imported there with stubbed `nvdiffrast` / `imageio` / `cv2`):
`apply_hdr_preprocessing` :263-286, `latlong_to_cubemap_official` :161-206, `latlong_vec` :320-338, `rgb2srgb_official` /
`reinhard_official` / `hdr_mapping_official` :109-140, the bilinear resize of `tonemap_image_direct` :493-497.
**PARITY UNPINNED for one stage:** the cube-map fetch of `render_projection_from_panorama` is
`nvdiffrast.torch.texture(filter_mode='linear', boundary_mode='cube')` (:446-447) — nvdiffrast (requirements.txt, no pinned
version) is not installed or vendored; `cube_texture_linear` restates its documented behaviour (bilinear over texel centres,
seamless across face edges) and is checked only for self-consistency (a cube map sampled back at its own texel-centre
directions is reproduced exactly).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def rgb2srgb(rgb: torch.Tensor) -> torch.Tensor:
    return torch.where(rgb <= 0.0031308, 12.92 * rgb, 1.055 * torch.pow(torch.clamp(rgb, 1e-8, 1.0), 1.0 / 2.4) - 0.055)


def reinhard(x: torch.Tensor, max_point: float = 16.0) -> torch.Tensor:
    return x / (x + 1.0) * max_point


def hdr_mapping(env_hdr: torch.Tensor, log_scale: float = 10000.0) -> Dict[str, torch.Tensor]:
    env_ev0 = rgb2srgb(reinhard(env_hdr, max_point=16.0).clamp(0, 1))
    env_log = rgb2srgb(torch.log1p(env_hdr) / np.log1p(log_scale)).clamp(0, 1)
    return {"env_hdr": env_hdr, "env_ev0": env_ev0, "env_log": env_log}


def apply_hdr_preprocessing(latlong_img: torch.Tensor, env_brightness: float, env_flip: bool, env_rot: float) -> torch.Tensor:
    latlong_img = latlong_img.clone()
    if env_brightness != 1.0:
        latlong_img *= env_brightness
    latlong_img = torch.nan_to_num(latlong_img, nan=0.0, posinf=65504.0, neginf=0.0).clamp(0.0, 65504.0)
    if env_flip:
        latlong_img = torch.flip(latlong_img, dims=[1])
    if env_rot != 0:
        latlong_img = torch.roll(latlong_img, shifts=int(latlong_img.shape[1] * env_rot / 360), dims=1)
    return latlong_img


def cube_to_dir(s: int, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    one = torch.ones_like(x)
    return [torch.stack([one, -y, -x], -1), torch.stack([-one, -y, x], -1), torch.stack([x, one, y], -1),
            torch.stack([x, -one, -y], -1), torch.stack([x, -y, one], -1), torch.stack([-x, -y, -one], -1)][s]


def latlong_to_cubemap(latlong_map: torch.Tensor, res: List[int]) -> torch.Tensor:
    device = latlong_map.device
    cubemap = torch.zeros(6, res[0], res[1], latlong_map.shape[-1], dtype=torch.float32, device=device)
    for s in range(6):
        gy, gx = torch.meshgrid(torch.linspace(-1.0 + 1.0 / res[0], 1.0 - 1.0 / res[0], res[0], device=device),
                                torch.linspace(-1.0 + 1.0 / res[1], 1.0 - 1.0 / res[1], res[1], device=device), indexing="ij")
        v = cube_to_dir(s, gx, gy)
        v = v / (torch.norm(v, dim=-1, keepdim=True) + 1e-8)
        tu = torch.atan2(v[..., 0:1], -v[..., 2:3]) / (2 * np.pi) + 0.5
        tv = torch.acos(torch.clamp(v[..., 1:2], min=-1, max=1)) / np.pi
        grid = (torch.cat((tu, tv), dim=-1) * 2.0 - 1.0).unsqueeze(0)
        sampled = F.grid_sample(latlong_map.permute(2, 0, 1).unsqueeze(0), grid, mode="bilinear", padding_mode="border", align_corners=False)
        cubemap[s, ...] = sampled.squeeze(0).permute(1, 2, 0)
    return cubemap


def latlong_vec(res: Tuple[int, int], device="cpu") -> torch.Tensor:
    H, W = res
    gy, gx = torch.meshgrid(torch.linspace(0.0 + 1.0 / H, 1.0 - 1.0 / H, H, device=device),
                            torch.linspace(-1.0 + 1.0 / W, 1.0 - 1.0 / W, W, device=device), indexing="ij")
    sintheta, costheta = torch.sin(gy * np.pi), torch.cos(gy * np.pi)
    sinphi, cosphi = torch.sin(gx * np.pi), torch.cos(gx * np.pi)
    return torch.stack((sintheta * sinphi, costheta, -sintheta * cosphi), dim=-1)


def dir_to_cube(d: torch.Tensor):
    """inverse of cube_to_dir: (face, x, y) of directions d [..., 3]"""
    dx, dy, dz = d.unbind(-1)
    ax, ay, az = dx.abs(), dy.abs(), dz.abs()
    is_x = (ax >= ay) & (ax >= az)
    is_y = ~is_x & (ay >= az)
    is_z = ~is_x & ~is_y
    s = torch.zeros_like(dx, dtype=torch.long)
    x, y = torch.zeros_like(dx), torch.zeros_like(dx)
    for cond, face, fx, fy, m in (
            (is_x & (dx > 0), 0, -dz, -dy, ax), (is_x & ~(dx > 0), 1, dz, -dy, ax),
            (is_y & (dy > 0), 2, dx, dz, ay), (is_y & ~(dy > 0), 3, dx, -dz, ay),
            (is_z & (dz > 0), 4, dx, -dy, az), (is_z & ~(dz > 0), 5, -dx, -dy, az)):
        s = torch.where(cond, torch.full_like(s, face), s)
        x = torch.where(cond, fx / m, x)
        y = torch.where(cond, fy / m, y)
    return s, x, y


def cube_texture_linear(cubemap: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """dr.texture(cubemap[None], d[None], filter_mode='linear', boundary_mode='cube')[0] restated (UNPINNED, see header)"""
    R = cubemap.shape[1]
    s, x, y = dir_to_cube(d)
    u, v = (x + 1) * 0.5 * R - 0.5, (y + 1) * 0.5 * R - 0.5
    x0, y0 = torch.floor(u), torch.floor(v)
    fx, fy = (u - x0).unsqueeze(-1), (v - y0).unsqueeze(-1)

    def texel(ix, iy):
        outside = (ix < 0) | (ix >= R) | (iy < 0) | (iy >= R)
        px, py = (2 * ix + 1) / R - 1, (2 * iy + 1) / R - 1
        dirs = torch.zeros(*ix.shape, 3, device=d.device)
        for f in range(6):
            dirs = torch.where((s == f).unsqueeze(-1), cube_to_dir(f, px, py), dirs)
        s2, nx, ny = dir_to_cube(dirs)
        jx = torch.floor((nx + 1) * 0.5 * R).clamp(0, R - 1)
        jy = torch.floor((ny + 1) * 0.5 * R).clamp(0, R - 1)
        fs = torch.where(outside, s2, s)
        fxi = torch.where(outside, jx, ix).long()
        fyi = torch.where(outside, jy, iy).long()
        return cubemap[fs, fyi, fxi]

    t00, t01, t10, t11 = texel(x0, y0), texel(x0 + 1, y0), texel(x0, y0 + 1), texel(x0 + 1, y0 + 1)
    return (t00 * (1 - fx) + t01 * fx) * (1 - fy) + (t10 * (1 - fx) + t11 * fx) * fy


def render_projection_from_panorama(latlong: torch.Tensor, resolution, env_brightness=1.0, env_flip=True, env_rot=180.0,
                                    cube_res: int = 512) -> Dict[str, torch.Tensor]:
    """reference :408-467 for a (He,We,3) tensor input: {'cubemap', 'env_ldr', 'env_log'} with (H,W,3) images"""
    H, W = resolution
    cubemap = latlong_to_cubemap(apply_hdr_preprocessing(latlong.float(), env_brightness, env_flip, env_rot), [cube_res, cube_res])
    vec_query = latlong_vec((H, W), device=latlong.device)          # c2w and y_rot are identities
    env_proj = torch.flip(cube_texture_linear(cubemap, -vec_query), dims=[0, 1])
    m = hdr_mapping(env_proj)
    return {"cubemap": cubemap, "env_ldr": m["env_ev0"], "env_log": m["env_log"]}


def tonemap_image_direct(img: torch.Tensor, resolution) -> Dict[str, torch.Tensor]:
    H, W = resolution
    if img.shape[:2] != (H, W):
        img = F.interpolate(img.permute(2, 0, 1).unsqueeze(0), size=(H, W), mode="bilinear", align_corners=False).squeeze(0).permute(1, 2, 0)
    m = hdr_mapping(img)
    return {"env_ldr": m["env_ev0"], "env_log": m["env_log"]}
