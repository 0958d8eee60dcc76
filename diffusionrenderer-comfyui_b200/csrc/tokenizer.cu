// HBM-bound kernels of the CV8x8x8 causal video tokenizer (diffusers.AutoencoderKLCosmos behind CleanVAE.py:44-60;
// SURVEY.md Appendix B): 3-D Haar patching and its inverse, the per-frame GroupNorm (+SiLU) apply pass, the row softmax
// and transpose used by the mid-block spatial attention, the causal temporal attention over <= 32 latent frames, and
// the planar <-> channels-last layout changes at the tokenizer boundary.  The convolutions are in conv.cu.
//
// Layouts: pixel-space video is planar bf16 [C][T][H][W] (the reference's BCTHW with B = 1); everything inside the
// tokenizer is channels-last per frame, bf16 [T][H][W][C].
//
// Roofline: HBM.  Algorithmic bytes per kernel are listed in DESIGN.md.
#include <math.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "postprocess.cuh"
#include "ptx.cuh"

namespace drb {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ===================================================================== Haar patching
// 4-point transform along one axis of the 4x4x4 block: index i = 2*i2 + i1 (i1: level-1 pair, i2: level-2 pair) ->
// j = b1 + 2*b2 (b1 / b2: low(0) or high(1) band of level 1 / level 2).  Symmetric, and H*H = 4*I.
__device__ __forceinline__ void had4(float& a0, float& a1, float& a2, float& a3) {
  const float s01 = a0 + a1, d01 = a0 - a1, s23 = a2 + a3, d23 = a2 - a3;
  a0 = s01 + s23;   // b1 = 0, b2 = 0
  a1 = d01 + d23;   // b1 = 1, b2 = 0
  a2 = s01 - s23;   // b1 = 0, b2 = 1
  a3 = d01 - d23;   // b1 = 1, b2 = 1
}
__device__ __forceinline__ void had4x4x4(float (&v)[4][4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      had4(v[a][b][0], v[a][b][1], v[a][b][2], v[a][b][3]);
    }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      had4(v[a][0][b], v[a][1][b], v[a][2][b], v[a][3][b]);
    }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      had4(v[0][a][b], v[1][a][b], v[2][a][b], v[3][a][b]);
    }
}
// sub-band channel of coefficient (jt, jh, jw) for input channel c of C: level-2 band * 8C + level-1 band * C + c, band
// index = (time, height, width) bits, time first (lll, llh, lhl, lhh, hll, ...).
__device__ __forceinline__ int haar_channel(int jt, int jh, int jw, int c, int C) {
  const int b1 = (jt & 1) * 4 + (jh & 1) * 2 + (jw & 1);
  const int b2 = (jt >> 1) * 4 + (jh >> 1) * 2 + (jw >> 1);
  return b2 * 8 * C + b1 * C + c;
}

constexpr int kHaarPix = 64;        // output pixels (along w') per block
constexpr int kHaarMaxC = 4;

// x [C][T][H][W] -> out [Tp][H/4][W/4][64*C], Tp = (T + 3) / 4; frame p of the padded clip is x[max(p - 3, 0)].
__global__ void __launch_bounds__(kHaarPix)
haar_patch_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int T, int H, int W) {
  extern __shared__ __align__(16) uint8_t hsm[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(hsm);
  const int CC = 64 * C, pitch = CC + 4;          // +4 elements: 2-way instead of 32-way bank conflicts
  const int Wp = W >> 2, Hp = H >> 2;
  const int wp0 = blockIdx.x * kHaarPix, hp = blockIdx.y, tp = blockIdx.z;
  const int wp = wp0 + threadIdx.x;
  if (wp < Wp) {
    for (int c = 0; c < C; ++c) {
      float v[4][4][4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int ts = max(4 * tp + it - 3, 0);
#pragma unroll
        for (int ih = 0; ih < 4; ++ih) {
          const __nv_bfloat16* src = x + ((static_cast<int64_t>(c) * T + ts) * H + (4 * hp + ih)) * W + 4 * wp;
          const uint2 r = *reinterpret_cast<const uint2*>(src);
          v[it][ih][0] = bf16_lo(r.x);
          v[it][ih][1] = bf16_hi(r.x);
          v[it][ih][2] = bf16_lo(r.y);
          v[it][ih][3] = bf16_hi(r.y);
        }
      }
      had4x4x4(v);
      __nv_bfloat16* row = tile + threadIdx.x * pitch;
#pragma unroll
      for (int jt = 0; jt < 4; ++jt)
#pragma unroll
        for (int jh = 0; jh < 4; ++jh)
#pragma unroll
          for (int jw = 0; jw < 4; ++jw)
            row[haar_channel(jt, jh, jw, c, C)] = __float2bfloat16_rn(v[jt][jh][jw] * (1.0f / 64.0f));
    }
  }
  __syncthreads();
  const int npix = min(kHaarPix, Wp - wp0);
  __nv_bfloat16* dst = out + ((static_cast<int64_t>(tp) * Hp + hp) * Wp + wp0) * CC;
  const int words_per_row = CC / 4;               // uint2 = 4 bf16
  for (int i = threadIdx.x; i < npix * words_per_row; i += kHaarPix) {
    const int pr = i / words_per_row, pc = i - pr * words_per_row;
    reinterpret_cast<uint2*>(dst)[i] = *reinterpret_cast<const uint2*>(tile + pr * pitch + pc * 4);
  }
}

// in [Tp][Hp][Wp][64*C] -> out [C][T][4Hp][4Wp], T = 4*Tp - 3 (the first 3 reconstructed frames are dropped).
// kU8 (C == 3): the decode post-process rides on the store — instead of the planar bf16 video the kernel writes the uint8
// BTHWC frames of diffusion_renderer_pipeline.py:300-318 (same values: the pixel is rounded to bf16 first, as the planar
// store would), saving the 308 MB write + read of the video tensor per decoded clip.
template <bool kU8>
__global__ void __launch_bounds__(kHaarPix)
haar_unpatch_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ out_u8, int C,
                    int Tp, int Hp, int Wp, float scale, int normalize_normal) {
  extern __shared__ __align__(16) uint8_t hsm[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(hsm);
  const int CC = 64 * C, pitch = CC + 4;
  const int T = 4 * Tp - 3, H = 4 * Hp, W = 4 * Wp;
  const int wp0 = blockIdx.x * kHaarPix, hp = blockIdx.y, tp = blockIdx.z;
  const int npix = min(kHaarPix, Wp - wp0);
  const __nv_bfloat16* src = in + ((static_cast<int64_t>(tp) * Hp + hp) * Wp + wp0) * CC;
  const int words_per_row = CC / 4;
  for (int i = threadIdx.x; i < npix * words_per_row; i += kHaarPix) {
    const int pr = i / words_per_row, pc = i - pr * words_per_row;
    *reinterpret_cast<uint2*>(tile + pr * pitch + pc * 4) = reinterpret_cast<const uint2*>(src)[i];
  }
  __syncthreads();
  const int wp = wp0 + threadIdx.x;
  if (wp >= Wp) return;
  const __nv_bfloat16* row = tile + threadIdx.x * pitch;
  // kU8: this thread's reconstructed pixels, [c][it][ih][iw] as bf16, parked in a private row behind the input tiles
  __nv_bfloat16* stash = tile + kHaarPix * pitch + threadIdx.x * pitch;
  for (int c = 0; c < C; ++c) {
    float v[4][4][4];
#pragma unroll
    for (int jt = 0; jt < 4; ++jt)
#pragma unroll
      for (int jh = 0; jh < 4; ++jh)
#pragma unroll
        for (int jw = 0; jw < 4; ++jw) v[jt][jh][jw] = __bfloat162float(row[haar_channel(jt, jh, jw, c, C)]);
    had4x4x4(v);   // inverse of (H x H x H) / 64 is H x H x H
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int t = 4 * tp + it - 3;
      if (t < 0) continue;
#pragma unroll
      for (int ih = 0; ih < 4; ++ih) {
        const uint2 packed = make_uint2(pack_bf16x2(v[it][ih][0], v[it][ih][1]), pack_bf16x2(v[it][ih][2], v[it][ih][3]));
        if constexpr (kU8) {
          *reinterpret_cast<uint2*>(stash + c * 64 + (it * 4 + ih) * 4) = packed;
        } else {
          __nv_bfloat16* dst = out + ((static_cast<int64_t>(c) * T + t) * H + (4 * hp + ih)) * W + 4 * wp;
          *reinterpret_cast<uint2*>(dst) = packed;
        }
      }
    }
  }
  if constexpr (kU8) {
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
      const int t = 4 * tp + it - 3;
      if (t < 0) continue;
#pragma unroll
      for (int ih = 0; ih < 4; ++ih) {
        uint8_t bytes[12];
#pragma unroll
        for (int iw = 0; iw < 4; ++iw) {
          float px[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) px[c] = bf16_round(__bfloat162float(stash[c * 64 + (it * 4 + ih) * 4 + iw]) * scale);
          uint8_t o[3];
          postprocess_pixel(px, normalize_normal, o);
          bytes[iw * 3] = o[0];
          bytes[iw * 3 + 1] = o[1];
          bytes[iw * 3 + 2] = o[2];
        }
        // 4 pixels x RGB = 12 contiguous bytes of the [T][H][W][3] frame (4-byte aligned: the pixel offset is a multiple of 4)
        uint32_t* dst = reinterpret_cast<uint32_t*>(out_u8 + ((static_cast<int64_t>(t) * H + (4 * hp + ih)) * W + 4 * wp) * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          dst[k] = bytes[4 * k] | (bytes[4 * k + 1] << 8) | (bytes[4 * k + 2] << 16) | (static_cast<uint32_t>(bytes[4 * k + 3]) << 24);
      }
    }
  }
}

// ===================================================================== per-frame GroupNorm(1 group) statistics / apply
// stats[t] = (sum, sum of squares) over the H*W*C values of frame t — the stand-alone form of what conv.cu's epilogue
// accumulates.  One block per (frame, slice); double atomics.
__global__ void __launch_bounds__(256)
frame_stats_kernel(const __nv_bfloat16* __restrict__ x, double* __restrict__ stats, int64_t per_frame) {
  const int t = blockIdx.y;
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(t) * per_frame);
  const int64_t n8 = per_frame >> 3;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * 256) {
    const uint4 v = src[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
      s1 += a + b;
      s2 += a * a + b * b;
    }
  }
  s1 = warp_sum_f(s1);
  s2 = warp_sum_f(s2);
  __shared__ float sh[2][8];
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s1;
    sh[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) {
      a += sh[0][i];
      b += sh[1][i];
    }
    atomicAdd(&stats[2 * t], a);
    atomicAdd(&stats[2 * t + 1], b);
  }
}

// round-to-nearest-even to bf16 precision with integer ALU ops, FINITE inputs only: the conversion instruction shares the
// XU pipe with the exponential and the reciprocal of SiLU, which ncu shows 79 % busy in this kernel.  (The GPU's canonical
// NaN, 0x7fffffff, would be carried into -0 by the rounding add: the caller re-injects non-finite inputs, see below.)
__device__ __forceinline__ float bf16_round_alu(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x7fffu + ((u >> 16) & 1u);
  return __uint_as_float(u & 0xffff0000u);
}

// out = act(bf16((x - mean_t) * rstd_t * gamma[c] + beta[c])), act = SiLU (rounded again) or identity; eps 1e-6;
// mean_t and rstd_t rounded to bf16 (see below).
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const double* __restrict__ stats,
                       const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta, int64_t per_frame, int C,
                       int silu) {
  const int t = blockIdx.y;
  const double n = static_cast<double>(per_frame);
  const double mean_d = stats[2 * t] / n;
  double var_d = stats[2 * t + 1] / n - mean_d * mean_d;
  if (var_d < 0.0) var_d = 0.0;
  // PyTorch's CUDA GroupNorm keeps the per-group mean and rstd in the *input* dtype when the affine parameters share it
  // (native/cuda/group_norm_kernel.cu: RowwiseMoments writes T, ComputeFusedParams reads T), so the reference pipeline in
  // bf16 normalises with bf16-rounded statistics; reproduced here (verified bit-equal on B200, tests/test_tokenizer_gpu.py).
  const float mean = bf16_round(static_cast<float>(mean_d));
  const float rstd = bf16_round(rsqrtf(static_cast<float>(var_d) + 1e-6f));
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(t) * per_frame);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(t) * per_frame);
  const int n8 = static_cast<int>(per_frame >> 3);     // per_frame < 2^34 is checked by the host
  const int c8 = C >> 3;
  const bool c8_pow2 = (c8 & (c8 - 1)) == 0;
#pragma unroll 2
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n8; i += gridDim.x * 256) {
    const int cg = (c8_pow2 ? (i & (c8 - 1)) : (i % c8)) * 8;   // the runtime modulus costs a reciprocal + two conversions (XU)
    const uint4 v = src[i];
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(gamma + cg));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(beta + cg));
    const uint32_t vw[4] = {v.x, v.y, v.z, v.w}, gw[4] = {g.x, g.y, g.z, g.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y0 = (bf16_lo(vw[j]) - mean) * rstd * bf16_lo(gw[j]) + bf16_lo(bw[j]);
      float y1 = (bf16_hi(vw[j]) - mean) * rstd * bf16_hi(gw[j]) + bf16_hi(bw[j]);
      if (silu) {   // x * sigmoid(x): exp and reciprocal are both MUFU ops (the kernel's floor); no full-precision division
        const float r0 = bf16_round_alu(y0), r1 = bf16_round_alu(y1);      // the reference rounds between the norm and SiLU
        // + 0 * y: nothing for finite y (the zero carries y's sign), NaN for a NaN / Inf y, which the integer rounding lost
        y0 = fmaf(y0, 0.0f, __fdividef(r0, 1.0f + __expf(-r0)));
        y1 = fmaf(y1, 0.0f, __fdividef(r1, 1.0f + __expf(-r1)));
      }
      o[j] = pack_bf16x2(y0, y1);   // the conversion rounds (once, for the identity activation) and keeps NaN / Inf
    }
    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ===================================================================== row softmax (in place, bf16 storage, fp32 math)
// s[r, 0:cols] <- softmax(scale * s[r, 0:cols]); s[r, cols:ld] <- 0 (K padding of the P.V GEMM).  One block per row:
// rows of up to 16384 columns live in registers (one read, one write); longer rows take three passes through L1/L2.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(__nv_bfloat16* __restrict__ s, int64_t ld, int cols, float scale_log2e) {
  __nv_bfloat16* row = s + static_cast<int64_t>(blockIdx.x) * ld;
  uint4* row4 = reinterpret_cast<uint4*>(row);
  const int n8 = static_cast<int>(ld >> 3);
  __shared__ float red[8];
  __shared__ float bcast;
  auto block_reduce = [&](float v, bool is_max) {
    v = is_max ? warp_max_f(v) : warp_sum_f(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float r = red[0];
      for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
      bcast = r;
    }
    __syncthreads();
    return bcast;
  };
  if (n8 <= 256 * 8) {
    // the whole row fits in registers (<= 16384 columns: 8 vectors per thread): one read, one write
    uint4 v[8];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n8) {
        v[k] = row4[i];
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (i * 8 + 2 * j < cols) m = fmaxf(m, bf16_lo(w[j]));
          if (i * 8 + 2 * j + 1 < cols) m = fmaxf(m, bf16_hi(w[j]));
        }
      }
    }
    m = block_reduce(m, true) * scale_log2e;
    float e[8][8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n8) {
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          e[k][2 * j] = (i * 8 + 2 * j < cols) ? exp2f(bf16_lo(w[j]) * scale_log2e - m) : 0.f;
          e[k][2 * j + 1] = (i * 8 + 2 * j + 1 < cols) ? exp2f(bf16_hi(w[j]) * scale_log2e - m) : 0.f;
          sum += e[k][2 * j] + e[k][2 * j + 1];
        }
      }
    }
    const float inv = 1.0f / block_reduce(sum, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n8)
        row4[i] = make_uint4(pack_bf16x2(e[k][0] * inv, e[k][1] * inv), pack_bf16x2(e[k][2] * inv, e[k][3] * inv),
                             pack_bf16x2(e[k][4] * inv, e[k][5] * inv), pack_bf16x2(e[k][6] * inv, e[k][7] * inv));
    }
    return;
  }
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n8; i += 256) {
    const uint4 v = row4[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i * 8 + 2 * j < cols) m = fmaxf(m, bf16_lo(w[j]));
      if (i * 8 + 2 * j + 1 < cols) m = fmaxf(m, bf16_hi(w[j]));
    }
  }
  m = block_reduce(m, true) * scale_log2e;
  float sum = 0.f;
  for (int i = threadIdx.x; i < n8; i += 256) {
    const uint4 v = row4[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i * 8 + 2 * j < cols) sum += exp2f(bf16_lo(w[j]) * scale_log2e - m);
      if (i * 8 + 2 * j + 1 < cols) sum += exp2f(bf16_hi(w[j]) * scale_log2e - m);
    }
  }
  const float inv = 1.0f / block_reduce(sum, false);
  for (int i = threadIdx.x; i < n8; i += 256) {
    const uint4 v = row4[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = (i * 8 + 2 * j < cols) ? exp2f(bf16_lo(w[j]) * scale_log2e - m) * inv : 0.f;
      const float b = (i * 8 + 2 * j + 1 < cols) ? exp2f(bf16_hi(w[j]) * scale_log2e - m) * inv : 0.f;
      o[j] = pack_bf16x2(a, b);
    }
    row4[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ===================================================================== transpose: in [rows][cols] -> out [cols][ld_out], pad zeroed
__global__ void __launch_bounds__(256)
transpose_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in, __nv_bfloat16* __restrict__ out, int64_t ld_out, int rows,
                 int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[static_cast<int64_t>(r) * ld_in + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < ld_out) out[static_cast<int64_t>(c) * ld_out + r] = tile[tx][i];   // r in [rows, ld_out) gets the zeros
  }
}

// ===================================================================== causal temporal attention
// qkv [T][HW][3C] (q | k | v), one head of dim C; out[t, p, :] = sum_{u <= t} softmax_u(q_t . k_u / sqrt(C)) v_u.
// One warp per pixel, online softmax over u; lane l owns channels {2l + 64k, 2l + 64k + 1}.
constexpr int kTAttnMaxK = 8;   // C <= 512

__global__ void __launch_bounds__(128)
temporal_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int64_t HW, int C,
                          float scale_log2e) {
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (pix >= HW) return;
  const int lane = threadIdx.x & 31;
  const int nk = C >> 6;
  const int64_t row_pitch = 3LL * C, frame_pitch = HW * row_pitch;
  const __nv_bfloat16* base = qkv + pix * row_pitch + 2 * lane;
  for (int t = 0; t < T; ++t) {
    float q[2 * kTAttnMaxK], acc[2 * kTAttnMaxK];
#pragma unroll
    for (int k = 0; k < kTAttnMaxK; ++k)
      if (k < nk) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(base + t * frame_pitch + 64 * k);
        q[2 * k] = bf16_lo(w);
        q[2 * k + 1] = bf16_hi(w);
        acc[2 * k] = acc[2 * k + 1] = 0.f;
      }
    float m = -INFINITY, l = 0.f;
    for (int u = 0; u <= t; ++u) {
      const __nv_bfloat16* kr = base + u * frame_pitch + C;
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < kTAttnMaxK; ++k)
        if (k < nk) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(kr + 64 * k);
          d += q[2 * k] * bf16_lo(w) + q[2 * k + 1] * bf16_hi(w);
        }
      const float s = warp_sum_f(d) * scale_log2e;
      const float m_new = fmaxf(m, s);
      const float corr = exp2f(m - m_new), p = exp2f(s - m_new);
      l = l * corr + p;
      m = m_new;
      const __nv_bfloat16* vr = kr + C;
#pragma unroll
      for (int k = 0; k < kTAttnMaxK; ++k)
        if (k < nk) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(vr + 64 * k);
          acc[2 * k] = acc[2 * k] * corr + p * bf16_lo(w);
          acc[2 * k + 1] = acc[2 * k + 1] * corr + p * bf16_hi(w);
        }
    }
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = out + (static_cast<int64_t>(t) * HW + pix) * C + 2 * lane;
#pragma unroll
    for (int k = 0; k < kTAttnMaxK; ++k)
      if (k < nk) *reinterpret_cast<uint32_t*>(orow + 64 * k) = pack_bf16x2(acc[2 * k] * inv, acc[2 * k + 1] * inv);
  }
}

// ===================================================================== layout changes at the tokenizer boundary
// x [C][T][H][W] -> out [T][H][W][Cpad] (channels >= C zero), values scaled by `scale` in fp32 before rounding.
__global__ void __launch_bounds__(256)
planar_to_cl_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int Cpad, int64_t thw, float scale) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;   // one thread per (pixel, group of 8 channels)
  const int groups = Cpad >> 3;
  const int64_t pix = i / groups;
  const int g = static_cast<int>(i - pix * groups);
  if (pix >= thw) return;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    v[j] = c < C ? __bfloat162float(x[static_cast<int64_t>(c) * thw + pix]) * scale : 0.f;
  }
  *reinterpret_cast<uint4*>(out + pix * Cpad + g * 8) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// in [T][H][W][Cpad] -> out [C][T][H][W] (first C channels), scaled.
__global__ void __launch_bounds__(256)
cl_to_planar_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int Cpad, int64_t thw, float scale) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;   // one thread per (channel, pixel), pixel fastest
  if (i >= thw * C) return;
  const int c = static_cast<int>(i / thw);
  const int64_t pix = i - static_cast<int64_t>(c) * thw;
  out[i] = __float2bfloat16_rn(__bfloat162float(in[pix * Cpad + c]) * scale);
}

}  // namespace
}  // namespace drb

using namespace drb;
#define STREAM static_cast<cudaStream_t>(stream)
#define BF(p) static_cast<__nv_bfloat16*>(p)
#define CBF(p) static_cast<const __nv_bfloat16*>(p)

extern "C" int drb_haar_patch(const void* x, void* out, int C, int T, int H, int W, void* stream) {
  DRB_REQUIRE(x && out, "null pointer");
  DRB_REQUIRE(C >= 1 && C <= kHaarMaxC, "1..4 input channels");
  DRB_REQUIRE(T >= 1 && (T - 1) % 4 == 0, "frame count must be 1 + 4k (first frame is repeated 4 times)");
  DRB_REQUIRE(H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, "height and width must be multiples of 4");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "pointers must be 8-byte aligned");
  const int Tp = (T + 3) / 4, Hp = H / 4, Wp = W / 4;
  DRB_REQUIRE(Hp <= 65535 && Tp <= 65535, "clip too large");
  const size_t smem = static_cast<size_t>(kHaarPix) * (64 * C + 4) * 2;
  static drb::DeviceOnce configured;   // per device: the attribute belongs to the device's context
  {
    const int rc = drb::device_once(configured, [] {
      return drb::check_cuda(cudaFuncSetAttribute(haar_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaarPix * (64 * kHaarMaxC + 4) * 2),
                             "haar_patch_kernel smem");
    });
    if (rc) return rc;
  }
  haar_patch_kernel<<<dim3((Wp + kHaarPix - 1) / kHaarPix, Hp, Tp), kHaarPix, smem, STREAM>>>(CBF(x), BF(out), C, T, H, W);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

static int haar_unpatch_launch(const void* in, void* out, void* out_u8, int C, int Tp, int Hp, int Wp, float scale,
                               int normalize_normal, void* stream) {
  DRB_REQUIRE(in && (out || out_u8), "null pointer");
  DRB_REQUIRE(C >= 1 && C <= kHaarMaxC, "1..4 output channels");
  DRB_REQUIRE(Tp >= 1 && Hp >= 1 && Wp >= 1 && Hp <= 65535 && Tp <= 65535, "bad dims");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(out_u8) & 3) == 0, "pointers must be 8-byte (uint8 output: 4-byte) aligned");
  const bool u8 = out_u8 != nullptr;
  DRB_REQUIRE(!u8 || C == 3, "the fused uint8 post-process needs a 3-channel video");
  // input tiles, plus (uint8 form) one private row per thread for the reconstructed pixels
  const size_t smem = static_cast<size_t>(kHaarPix) * (64 * C + 4) * 2 * (u8 ? 2 : 1);
  static drb::DeviceOnce configured;   // per device: the attribute belongs to the device's context
  {
    const int rc = drb::device_once(configured, [] {
      const int bytes = kHaarPix * (64 * kHaarMaxC + 4) * 2;
      int r = drb::check_cuda(cudaFuncSetAttribute(haar_unpatch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes),
                              "haar_unpatch_kernel smem");
      if (r) return r;
      return drb::check_cuda(cudaFuncSetAttribute(haar_unpatch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bytes),
                             "haar_unpatch_kernel<u8> smem");
    });
    if (rc) return rc;
  }
  const dim3 grid((Wp + kHaarPix - 1) / kHaarPix, Hp, Tp);
  if (u8)
    haar_unpatch_kernel<true><<<grid, kHaarPix, smem, STREAM>>>(CBF(in), nullptr, static_cast<uint8_t*>(out_u8), C, Tp, Hp, Wp, scale,
                                                                normalize_normal);
  else
    haar_unpatch_kernel<false><<<grid, kHaarPix, smem, STREAM>>>(CBF(in), BF(out), nullptr, C, Tp, Hp, Wp, 1.0f, 0);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_haar_unpatch(const void* in, void* out, int C, int Tp, int Hp, int Wp, void* stream) {
  DRB_REQUIRE(out != nullptr, "null pointer");
  return haar_unpatch_launch(in, out, nullptr, C, Tp, Hp, Wp, 1.0f, 0, stream);
}

extern "C" int drb_haar_unpatch_u8(const void* in, void* out_u8, int Tp, int Hp, int Wp, int normalize_normal, void* stream) {
  DRB_REQUIRE(out_u8 != nullptr, "null pointer");
  return haar_unpatch_launch(in, nullptr, out_u8, 3, Tp, Hp, Wp, 1.0f, normalize_normal, stream);
}

static int frame_grid_x(int64_t per_frame, int T) {
  int64_t want = (per_frame / 8 + 256 * 8 - 1) / (256 * 8);       // ~8 vectors per thread
  const int64_t cap = (static_cast<int64_t>(num_sms()) * 16 + T - 1) / T;
  if (want > cap) want = cap;
  return want < 1 ? 1 : static_cast<int>(want);
}

extern "C" int drb_frame_stats_cl(const void* x, double* stats, int T, int64_t per_frame, void* stream) {
  DRB_REQUIRE(x && stats, "null pointer");
  DRB_REQUIRE(T > 0 && T <= 65535 && per_frame > 0 && per_frame % 8 == 0, "per-frame element count must be a positive multiple of 8");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  DRB_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * T, STREAM));
  frame_stats_kernel<<<dim3(frame_grid_x(per_frame, T), T), 256, 0, STREAM>>>(CBF(x), stats, per_frame);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_groupnorm_apply_cl(const void* x, void* out, const double* stats, const void* gamma, const void* beta, int T,
                                      int64_t hw, int C, int silu, void* stream) {
  DRB_REQUIRE(x && out && stats && gamma && beta, "null pointer");
  DRB_REQUIRE(T > 0 && T <= 65535 && hw > 0 && C > 0 && C % 8 == 0, "C must be a multiple of 8");
  DRB_REQUIRE(hw * C < (1LL << 34), "frame too large");
  DRB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "pointers must be 16-byte aligned");
  const int64_t per_frame = hw * C;
  groupnorm_apply_kernel<<<dim3(frame_grid_x(per_frame, T), T), 256, 0, STREAM>>>(CBF(x), BF(out), stats, CBF(gamma), CBF(beta),
                                                                                  per_frame, C, silu);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_softmax_rows(void* s, int64_t ld, int rows, int cols, float scale, void* stream) {
  DRB_REQUIRE(s != nullptr, "null pointer");
  DRB_REQUIRE(rows > 0 && cols > 0 && ld >= cols && ld % 8 == 0, "row pitch must be a multiple of 8 and >= cols");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(s) & 15) == 0, "s must be 16-byte aligned");
  softmax_rows_kernel<<<rows, 256, 0, STREAM>>>(BF(s), ld, cols, scale * 1.4426950408889634f);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols, void* stream) {
  DRB_REQUIRE(in && out, "null pointer");
  DRB_REQUIRE(rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "bad pitches");
  const int rcover = static_cast<int>((ld_out + 31) / 32);
  DRB_REQUIRE(rcover <= 65535, "too many rows");
  transpose_kernel<<<dim3((cols + 31) / 32, rcover), 256, 0, STREAM>>>(CBF(in), ld_in, BF(out), ld_out, rows, cols);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_temporal_attention_cl(const void* qkv, void* out, int T, int64_t hw, int C, void* stream) {
  DRB_REQUIRE(qkv && out, "null pointer");
  DRB_REQUIRE(T > 0 && hw > 0, "bad dims");
  DRB_REQUIRE(C % 64 == 0 && C <= 64 * kTAttnMaxK, "C must be a multiple of 64, at most 512");
  temporal_attention_kernel<<<static_cast<unsigned>((hw + 3) / 4), 128, 0, STREAM>>>(CBF(qkv), BF(out), T, hw, C,
                                                                                    1.4426950408889634f / sqrtf(static_cast<float>(C)));
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_planar_to_cl(const void* x, void* out, int C, int Cpad, int64_t thw, float scale, void* stream) {
  DRB_REQUIRE(x && out, "null pointer");
  DRB_REQUIRE(C > 0 && Cpad >= C && Cpad % 8 == 0 && thw > 0, "Cpad must be a multiple of 8 and >= C");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  const int64_t n = thw * (Cpad / 8);
  planar_to_cl_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, STREAM>>>(CBF(x), BF(out), C, Cpad, thw, scale);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

// Per-chunk latent statistics of the chunking tokenizer (pretrained_vae.py:131-152): one (mean, std) per row of `hw` values.
//   mode 0 (after encode):  out = bf16(bf16(x - mean) / std)        (:142)
//   mode 1 (before decode): out = bf16(bf16(x * std) + mean)        (:150)   — both in the bf16 arithmetic of the reference
namespace drb {
namespace {
__global__ void __launch_bounds__(256)
latent_normalize_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ mean,
                        const __nv_bfloat16* __restrict__ stdv, __nv_bfloat16* __restrict__ out, int64_t hw, int mode) {
  const int row = blockIdx.y;
  const float m = __bfloat162float(mean[row]), s = __bfloat162float(stdv[row]);
  const __nv_bfloat16* xr = x + static_cast<int64_t>(row) * hw;
  __nv_bfloat16* orow = out + static_cast<int64_t>(row) * hw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < hw; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = __bfloat162float(xr[i]);
    const float o = mode == 0 ? __bfloat162float(__float2bfloat16_rn(v - m)) / s : __bfloat162float(__float2bfloat16_rn(v * s)) + m;
    orow[i] = __float2bfloat16_rn(o);
  }
}
}  // namespace
}  // namespace drb

extern "C" int drb_latent_normalize(const void* x, const void* mean, const void* stdv, void* out, int rows, int64_t hw, int mode,
                                    void* stream) {
  DRB_REQUIRE(x && mean && stdv && out, "null pointer");
  DRB_REQUIRE(rows > 0 && rows <= 65535 && hw > 0 && (mode == 0 || mode == 1), "bad sizes or mode");
  int gx = static_cast<int>((hw + 255) / 256);
  if (gx > 64) gx = 64;
  latent_normalize_kernel<<<dim3(gx, rows), 256, 0, STREAM>>>(CBF(x), CBF(mean), CBF(stdv), BF(out), hw, mode);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_cl_to_planar(const void* in, void* out, int C, int Cpad, int64_t thw, float scale, void* stream) {
  DRB_REQUIRE(in && out, "null pointer");
  DRB_REQUIRE(C > 0 && Cpad >= C && thw > 0, "Cpad must be >= C");
  const int64_t n = thw * C;
  cl_to_planar_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, STREAM>>>(CBF(in), BF(out), C, Cpad, thw, scale);
  DRB_CUDA(cudaGetLastError());
  return 0;
}
