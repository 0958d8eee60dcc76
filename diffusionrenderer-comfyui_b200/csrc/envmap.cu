// Environment-map conditioning of the forward renderer on the device, without nvdiffrast.
//
// Follows the reference preprocess_envmap.py: apply_hdr_preprocessing (:263-286: brightness, NaN/Inf clean-up, clamp,
// horizontal flip, roll) fused into latlong_to_cubemap_official (:161-206: 512x512 cube faces by bilinear grid_sample,
// border padding, align_corners=False); render_projection_from_panorama (:408-467: every output pixel of an (H, W)
// lat-long grid looks the cube map up in direction -latlong_vec, linear filter, result flipped in both axes) fused with
// hdr_mapping_official (:119-140: Reinhard + sRGB "ev0" image and log-encoded image); tonemap_image_direct (:469-526:
// bilinear resize + the same tone mapping) for pre-rendered light probes.
// The cube-map fetch restates nvdiffrast's dr.texture(filter_mode='linear', boundary_mode='cube'): bilinear over texel
// centres, taps that fall off a face are taken from the neighbouring face through their 3-D direction.
// fp32 throughout (HDR radiance); a few MB per call — latency-, not bandwidth-bound.
#include <math.h>

#include "../../include/drb200.h"
#include "common.cuh"

namespace drb {
namespace {

constexpr float kPi = 3.14159265358979323846f;

__device__ __forceinline__ float srgb(float x) {   // rgb2srgb_official (:109-113)
  return x <= 0.0031308f ? 12.92f * x : 1.055f * powf(fminf(fmaxf(x, 1e-8f), 1.0f), 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ void tone_map(float x, float& ldr, float& lg) {   // hdr_mapping_official (:119-140)
  const float r = fminf(fmaxf(x / (x + 1.0f) * 16.0f, 0.0f), 1.0f);
  ldr = srgb(r);
  lg = fminf(fmaxf(srgb(log1pf(x) / 9.210440366976517f /* log1p(1e4) */), 0.0f), 1.0f);
}
// one preprocessed panorama texel (:263-286): brightness, nan -> 0, +inf -> 65504, clamp, flip, roll
__device__ __forceinline__ float pano(const float* __restrict__ src, int We, int y, int x, int c, float brightness, int flip, int roll) {
  int xs = (x - roll) % We;
  if (xs < 0) xs += We;
  if (flip) xs = We - 1 - xs;
  float v = src[(static_cast<int64_t>(y) * We + xs) * 3 + c] * brightness;
  if (isnan(v)) v = 0.0f;
  return fminf(fmaxf(v, 0.0f), 65504.0f);
}
__device__ __forceinline__ void cube_to_dir(int s, float x, float y, float& dx, float& dy, float& dz) {   // (:142-155)
  switch (s) {
    case 0: dx = 1.f; dy = -y; dz = -x; break;
    case 1: dx = -1.f; dy = -y; dz = x; break;
    case 2: dx = x; dy = 1.f; dz = y; break;
    case 3: dx = x; dy = -1.f; dz = -y; break;
    case 4: dx = x; dy = -y; dz = 1.f; break;
    default: dx = -x; dy = -y; dz = -1.f; break;
  }
}
// inverse of cube_to_dir: face of the dominant axis and the in-face coordinates in [-1, 1]
__device__ __forceinline__ void dir_to_cube(float dx, float dy, float dz, int& s, float& x, float& y) {
  const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  if (ax >= ay && ax >= az) {
    const float inv = 1.0f / ax;
    if (dx > 0) { s = 0; x = -dz * inv; y = -dy * inv; } else { s = 1; x = dz * inv; y = -dy * inv; }
  } else if (ay >= az) {
    const float inv = 1.0f / ay;
    if (dy > 0) { s = 2; x = dx * inv; y = dz * inv; } else { s = 3; x = dx * inv; y = -dz * inv; }
  } else {
    const float inv = 1.0f / az;
    if (dz > 0) { s = 4; x = dx * inv; y = -dy * inv; } else { s = 5; x = -dx * inv; y = -dy * inv; }
  }
}

__global__ void __launch_bounds__(256)
latlong_to_cubemap_kernel(const float* __restrict__ src, int He, int We, float brightness, int flip, int roll, float* __restrict__ cube,
                          int R) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 6 * R * R) return;
  const int ix = i % R, iy = (i / R) % R, s = i / (R * R);
  const float step = R > 1 ? (2.0f - 2.0f / R) / (R - 1) : 0.f;          // torch.linspace(-1 + 1/R, 1 - 1/R, R)
  const float gx = -1.0f + 1.0f / R + ix * step, gy = -1.0f + 1.0f / R + iy * step;
  float dx, dy, dz;
  cube_to_dir(s, gx, gy, dx, dy, dz);
  const float inv = 1.0f / (sqrtf(dx * dx + dy * dy + dz * dz) + 1e-8f);  // safe_normalize
  dx *= inv; dy *= inv; dz *= inv;
  const float tu = atan2f(dx, -dz) / (2.0f * kPi) + 0.5f;
  const float tv = acosf(fminf(fmaxf(dy, -1.0f), 1.0f)) / kPi;
  // grid_sample(bilinear, padding_mode='border', align_corners=False) at grid = 2t - 1
  const float px = fminf(fmaxf(tu * We - 0.5f, 0.0f), static_cast<float>(We - 1));
  const float py = fminf(fmaxf(tv * He - 0.5f, 0.0f), static_cast<float>(He - 1));
  const int x0 = static_cast<int>(floorf(px)), y0 = static_cast<int>(floorf(py));
  const int x1 = min(x0 + 1, We - 1), y1 = min(y0 + 1, He - 1);
  const float fx = px - x0, fy = py - y0;
  for (int c = 0; c < 3; ++c) {
    const float v00 = pano(src, We, y0, x0, c, brightness, flip, roll), v01 = pano(src, We, y0, x1, c, brightness, flip, roll);
    const float v10 = pano(src, We, y1, x0, c, brightness, flip, roll), v11 = pano(src, We, y1, x1, c, brightness, flip, roll);
    cube[static_cast<int64_t>(i) * 3 + c] = (v00 * (1.f - fx) + v01 * fx) * (1.f - fy) + (v10 * (1.f - fx) + v11 * fx) * fy;
  }
}

// texel (s, iy, ix) with ix / iy possibly one step outside the face: re-enter through the 3-D direction
__device__ __forceinline__ const float* cube_texel(const float* __restrict__ cube, int R, int s, int iy, int ix) {
  if (ix < 0 || ix >= R || iy < 0 || iy >= R) {
    const float x = (2.0f * ix + 1.0f) / R - 1.0f, y = (2.0f * iy + 1.0f) / R - 1.0f;
    float dx, dy, dz, nx, ny;
    cube_to_dir(s, x, y, dx, dy, dz);
    dir_to_cube(dx, dy, dz, s, nx, ny);
    ix = min(max(static_cast<int>(floorf((nx + 1.0f) * 0.5f * R)), 0), R - 1);
    iy = min(max(static_cast<int>(floorf((ny + 1.0f) * 0.5f * R)), 0), R - 1);
  }
  return cube + ((static_cast<int64_t>(s) * R + iy) * R + ix) * 3;
}

__global__ void __launch_bounds__(256)
project_kernel(const float* __restrict__ cube, int R, float* __restrict__ ldr, float* __restrict__ lg, int H, int W) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= H * W) return;
  const int w = i % W, h = i / W;
  const int hs = H - 1 - h, wsrc = W - 1 - w;                              // torch.flip(env_proj, dims=[0, 1]) (:448)
  // latlong_vec (:320-338): linspace(1/H, 1 - 1/H, H) x linspace(-1 + 1/W, 1 - 1/W, W)
  const float gy = 1.0f / H + (H > 1 ? hs * ((1.0f - 2.0f / H) / (H - 1)) : 0.f);
  const float gx = -1.0f + 1.0f / W + (W > 1 ? wsrc * ((2.0f - 2.0f / W) / (W - 1)) : 0.f);
  const float st = sinf(gy * kPi), ct = cosf(gy * kPi), sp = sinf(gx * kPi), cp = cosf(gx * kPi);
  const float qx = -(st * sp), qy = -ct, qz = st * cp;                     // query direction = -vec
  int s;
  float x, y;
  dir_to_cube(qx, qy, qz, s, x, y);
  const float u = (x + 1.0f) * 0.5f * R - 0.5f, v = (y + 1.0f) * 0.5f * R - 0.5f;
  const int x0 = static_cast<int>(floorf(u)), y0 = static_cast<int>(floorf(v));
  const float fx = u - x0, fy = v - y0;
  const float* t00 = cube_texel(cube, R, s, y0, x0);
  const float* t01 = cube_texel(cube, R, s, y0, x0 + 1);
  const float* t10 = cube_texel(cube, R, s, y0 + 1, x0);
  const float* t11 = cube_texel(cube, R, s, y0 + 1, x0 + 1);
  for (int c = 0; c < 3; ++c) {
    const float val = (t00[c] * (1.f - fx) + t01[c] * fx) * (1.f - fy) + (t10[c] * (1.f - fx) + t11[c] * fx) * fy;
    tone_map(val, ldr[static_cast<int64_t>(i) * 3 + c], lg[static_cast<int64_t>(i) * 3 + c]);
  }
}

// F.interpolate(bilinear, align_corners=False) to (H, W), then the tone mapping (:493-505)
__global__ void __launch_bounds__(256)
tonemap_resize_kernel(const float* __restrict__ src, int Hs, int Ws, float* __restrict__ ldr, float* __restrict__ lg, int H, int W) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= H * W) return;
  const int w = i % W, h = i / W;
  const float sy = fmaxf((h + 0.5f) * (static_cast<float>(Hs) / H) - 0.5f, 0.0f);
  const float sx = fmaxf((w + 0.5f) * (static_cast<float>(Ws) / W) - 0.5f, 0.0f);
  const int y0 = min(static_cast<int>(sy), Hs - 1), x0 = min(static_cast<int>(sx), Ws - 1);
  const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
  const float fy = sy - y0, fx = sx - x0;
  for (int c = 0; c < 3; ++c) {
    const float v00 = src[(static_cast<int64_t>(y0) * Ws + x0) * 3 + c], v01 = src[(static_cast<int64_t>(y0) * Ws + x1) * 3 + c];
    const float v10 = src[(static_cast<int64_t>(y1) * Ws + x0) * 3 + c], v11 = src[(static_cast<int64_t>(y1) * Ws + x1) * 3 + c];
    const float val = (v00 * (1.f - fx) + v01 * fx) * (1.f - fy) + (v10 * (1.f - fx) + v11 * fx) * fy;
    tone_map(val, ldr[static_cast<int64_t>(i) * 3 + c], lg[static_cast<int64_t>(i) * 3 + c]);
  }
}

}  // namespace
}  // namespace drb

using namespace drb;

extern "C" int drb_envmap_latlong_to_cubemap(const float* latlong, int He, int We, float brightness, int flip, int roll_px,
                                             float* cube, int R, void* stream) {
  DRB_REQUIRE(latlong && cube, "null pointer");
  DRB_REQUIRE(He > 0 && We > 0 && R > 0 && R <= 4096, "bad sizes");
  const int n = 6 * R * R;
  latlong_to_cubemap_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(latlong, He, We, brightness, flip, roll_px,
                                                                                          cube, R);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_envmap_project(const float* cube, int R, float* env_ldr, float* env_log, int H, int W, void* stream) {
  DRB_REQUIRE(cube && env_ldr && env_log, "null pointer");
  DRB_REQUIRE(R > 0 && H > 0 && W > 0 && static_cast<int64_t>(H) * W < (1LL << 31), "bad sizes");
  project_kernel<<<(H * W + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(cube, R, env_ldr, env_log, H, W);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_envmap_tonemap(const float* src, int Hs, int Ws, float* env_ldr, float* env_log, int H, int W, void* stream) {
  DRB_REQUIRE(src && env_ldr && env_log, "null pointer");
  DRB_REQUIRE(Hs > 0 && Ws > 0 && H > 0 && W > 0 && static_cast<int64_t>(H) * W < (1LL << 31), "bad sizes");
  tonemap_resize_kernel<<<(H * W + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, Hs, Ws, env_ldr, env_log, H, W);
  DRB_CUDA(cudaGetLastError());
  return 0;
}
