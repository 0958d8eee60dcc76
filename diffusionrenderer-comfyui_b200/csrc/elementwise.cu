// Fused, vectorised, HBM-bound kernels of the denoising path: AdaLN modulation (+ the degenerate cross-attention
// residual), per-head RMSNorm + 3-D RoPE on q/k, the sigma-only GEMVs (timestep embedding, AdaLN-LoRA, context
// projections), patchify / unpatchify with the EDM c_in scaling and Euler update, and the uint8 post-process.
// Each kernel rounds to bf16 exactly where the reference's chain of bf16 tensor ops does.
//
// Roofline: HBM.  Algorithmic bytes per kernel are listed in DESIGN.md.
#include <math.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "postprocess.cuh"
#include "ptx.cuh"

namespace drb {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// ===================================================================== AdaLN modulate (+ optional residual add)
// One warp per token row; the row lives in registers (D <= 4096: 16 chunks of 32 lanes x 8 bf16).
constexpr int kAdaMaxChunks = 16;
constexpr int kAdaWarps = 8;

__global__ void __launch_bounds__(kAdaWarps * 32, 3)
adaln_kernel(__nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ shift,
             const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ add_gate,
             const __nv_bfloat16* __restrict__ add_vec, int rows, int D) {
  const int row = blockIdx.x * kAdaWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int chunks = D >> 8;   // D / 256
  uint4 v[kAdaMaxChunks];
  __nv_bfloat16* xr = x + static_cast<int64_t>(row) * D;
#pragma unroll
  for (int c = 0; c < kAdaMaxChunks; ++c)
    if (c < chunks) v[c] = *reinterpret_cast<const uint4*>(xr + c * 256 + lane * 8);

  if (add_vec != nullptr) {
    // x <- bf16(x + bf16(gate * vec)): the whole cross-attention sub-block for a one-token context
#pragma unroll
    for (int c = 0; c < kAdaMaxChunks; ++c)
      if (c < chunks) {
        const uint4 g = __ldg(reinterpret_cast<const uint4*>(add_gate + c * 256 + lane * 8));
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(add_vec + c * 256 + lane * 8));
        uint32_t* pv = reinterpret_cast<uint32_t*>(&v[c]);
        const uint32_t* pg = reinterpret_cast<const uint32_t*>(&g);
        const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float t0 = bf16_round(bf16_lo(pg[i]) * bf16_lo(pa[i]));
          const float t1 = bf16_round(bf16_hi(pg[i]) * bf16_hi(pa[i]));
          pv[i] = pack_bf16x2(bf16_lo(pv[i]) + t0, bf16_hi(pv[i]) + t1);
        }
        *reinterpret_cast<uint4*>(xr + c * 256 + lane * 8) = v[c];
      }
  }

  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kAdaMaxChunks; ++c)
    if (c < chunks) {
      const uint32_t* pv = reinterpret_cast<const uint32_t*>(&v[c]);
#pragma unroll
      for (int i = 0; i < 4; ++i) sum += bf16_lo(pv[i]) + bf16_hi(pv[i]);
    }
  const float mean = warp_sum(sum) / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < kAdaMaxChunks; ++c)
    if (c < chunks) {
      const uint32_t* pv = reinterpret_cast<const uint32_t*>(&v[c]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = bf16_lo(pv[i]) - mean, b = bf16_hi(pv[i]) - mean;
        sq += a * a + b * b;
      }
    }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(D) + 1e-6f);

  __nv_bfloat16* orow = out + static_cast<int64_t>(row) * D;
#pragma unroll
  for (int c = 0; c < kAdaMaxChunks; ++c)
    if (c < chunks) {
      const uint4 sh = __ldg(reinterpret_cast<const uint4*>(shift + c * 256 + lane * 8));
      const uint4 sc = __ldg(reinterpret_cast<const uint4*>(scale + c * 256 + lane * 8));
      const uint32_t* pv = reinterpret_cast<const uint32_t*>(&v[c]);
      const uint32_t* psh = reinterpret_cast<const uint32_t*>(&sh);
      const uint32_t* psc = reinterpret_cast<const uint32_t*>(&sc);
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float n0 = bf16_round((bf16_lo(pv[i]) - mean) * rstd);
        const float n1 = bf16_round((bf16_hi(pv[i]) - mean) * rstd);
        const float s0 = bf16_round(1.0f + bf16_lo(psc[i]));
        const float s1 = bf16_round(1.0f + bf16_hi(psc[i]));
        const float m0 = bf16_round(n0 * s0);
        const float m1 = bf16_round(n1 * s1);
        o[i] = pack_bf16x2(m0 + bf16_lo(psh[i]), m1 + bf16_hi(psh[i]));
      }
      *reinterpret_cast<uint4*>(orow + c * 256 + lane * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ===================================================================== per-head RMSNorm + RoPE on q and k, in place
// One CTA per token row; 8 warps sweep the 2*H head slots (q heads then k heads); a lane owns 4 columns and its
// rotate_half partner sits in lane ^ 16.
__global__ void __launch_bounds__(256)
qk_norm_rope_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ld, const __nv_bfloat16* __restrict__ wq,
                    const __nv_bfloat16* __restrict__ wk, const __nv_bfloat16* __restrict__ cos_tab,
                    const __nv_bfloat16* __restrict__ sin_tab, int S, int H) {
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint2 cv = __ldg(reinterpret_cast<const uint2*>(cos_tab + static_cast<int64_t>(row) * 128 + lane * 4));
  const uint2 sv = __ldg(reinterpret_cast<const uint2*>(sin_tab + static_cast<int64_t>(row) * 128 + lane * 4));
  const float c[4] = {bf16_lo(cv.x), bf16_hi(cv.x), bf16_lo(cv.y), bf16_hi(cv.y)};
  const float s[4] = {bf16_lo(sv.x), bf16_hi(sv.x), bf16_lo(sv.y), bf16_hi(sv.y)};
  const uint2 wqv = __ldg(reinterpret_cast<const uint2*>(wq + lane * 4));
  const uint2 wkv = __ldg(reinterpret_cast<const uint2*>(wk + lane * 4));
  const float sign = lane < 16 ? -1.0f : 1.0f;   // rotate_half = cat(-x[64:], x[:64])
  __nv_bfloat16* base = qkv + static_cast<int64_t>(row) * ld;
  for (int slot = warp; slot < 2 * H; slot += 8) {
    const bool is_k = slot >= H;
    const uint2 wv = is_k ? wkv : wqv;
    const float w[4] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y)};
    uint2* ptr = reinterpret_cast<uint2*>(base + static_cast<int64_t>(slot) * 128 + lane * 4);   // k block follows q block
    const uint2 xv = *ptr;
    float x[4] = {bf16_lo(xv.x), bf16_hi(xv.x), bf16_lo(xv.y), bf16_hi(xv.y)};
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
    const float inv = rsqrtf(warp_sum(ss) * (1.0f / 128.0f) + 1e-6f);
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = bf16_round(x[i] * inv * w[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float partner = __shfl_xor_sync(0xffffffffu, x[i], 16);
      const float a = bf16_round(x[i] * c[i]);
      const float b = bf16_round(sign * partner * s[i]);
      o[i] = a + b;
    }
    *ptr = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
  }
}

// ===================================================================== GEMV (batched): y_b = bf16(W_b act(x_b)) (+ add)
// One warp per output row; x staged once per CTA in shared memory as fp32.
constexpr int kGemvWarps = 8;
constexpr int kGemvMaxK = 4096;

__global__ void __launch_bounds__(kGemvWarps * 32)
gemv_kernel(const __nv_bfloat16* __restrict__ W, int64_t ldw, int64_t w_bs, const __nv_bfloat16* __restrict__ x,
            int64_t x_bs, __nv_bfloat16* __restrict__ y, int64_t y_bs, const __nv_bfloat16* __restrict__ add,
            int64_t add_bs, int N, int K, int act) {
  __shared__ float xs[kGemvMaxK];
  const int b = blockIdx.y;
  const __nv_bfloat16* xb = x + b * x_bs;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float v = __bfloat162float(xb[k]);
    if (act == 1) v = bf16_round(silu_f(v));
    xs[k] = v;
  }
  __syncthreads();
  const int n = blockIdx.x * kGemvWarps + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* wr = W + b * w_bs + static_cast<int64_t>(n) * ldw;
  float acc = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wr + k));
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(&wv);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += bf16_lo(pw[i]) * xs[k + 2 * i] + bf16_hi(pw[i]) * xs[k + 2 * i + 1];
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float r = bf16_round(acc);
    if (add != nullptr) r = r + __bfloat162float(add[b * add_bs + n]);
    y[b * y_bs + n] = __float2bfloat16_rn(r);
  }
}

// ===================================================================== sigma embedding
__global__ void __launch_bounds__(1024)
sigma_embedding_kernel(const float* __restrict__ sigma, const __nv_bfloat16* __restrict__ w_aff,
                       __nv_bfloat16* __restrict__ e_out, __nv_bfloat16* __restrict__ emb_out, int D) {
  __shared__ float red[32];
  __shared__ float inv_s;
  const int half = D / 2;
  const float st = bf16_round(*sigma);   // the reference rounds sigma to bf16 BEFORE the sinusoid (:664)
  float ss = 0.f;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float w = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
    const float ang = st * w;
    const float c = bf16_round(cosf(ang)), s = bf16_round(sinf(ang));
    e_out[i] = __float2bfloat16_rn(c);
    e_out[half + i] = __float2bfloat16_rn(s);
    ss += c * c + s * s;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) inv_s = rsqrtf(v / static_cast<float>(D) + 1e-6f);
  }
  __syncthreads();
  const float inv = inv_s;
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    emb_out[i] = __float2bfloat16_rn(__bfloat162float(e_out[i]) * inv * __bfloat162float(w_aff[i]));
}

// ===================================================================== patchify
// tokens[s, (c0 + c) * 4 + m * 2 + n] = f(src[c, t, 2h + m, 2w + n]);  s = (t * Hp + h) * Wp + w
template <bool kScale>
__global__ void __launch_bounds__(256)
patchify_kernel(const __nv_bfloat16* __restrict__ src, const float* __restrict__ sigma, __nv_bfloat16* __restrict__ tok,
                int64_t ld_tok, int c0, int C, int T, int H, int W) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int64_t total = static_cast<int64_t>(T) * Hp * Wp * C;
  float cin = 1.0f;
  if (kScale) {
    const float sg = *sigma;
    cin = 1.0f / sqrtf(sg * sg + 0.25f);
  }
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const int64_t s = idx / C;
    const int w = static_cast<int>(s % Wp);
    const int h = static_cast<int>((s / Wp) % Hp);
    const int t = static_cast<int>(s / (static_cast<int64_t>(Wp) * Hp));
    const __nv_bfloat16* p = src + ((static_cast<int64_t>(c) * T + t) * H + 2 * h) * W + 2 * w;
    const uint32_t r0 = *reinterpret_cast<const uint32_t*>(p);
    const uint32_t r1 = *reinterpret_cast<const uint32_t*>(p + W);
    uint2 o;
    if (kScale) {
      o.x = pack_bf16x2(bf16_lo(r0) * cin, bf16_hi(r0) * cin);
      o.y = pack_bf16x2(bf16_lo(r1) * cin, bf16_hi(r1) * cin);
    } else {
      o.x = r0;
      o.y = r1;
    }
    *reinterpret_cast<uint2*>(tok + s * ld_tok + (c0 + c) * 4) = o;
  }
}

__global__ void __launch_bounds__(256)
patchify_fill_kernel(__nv_bfloat16* __restrict__ tok, int64_t ld_tok, int64_t S, int ones_channel, int zero_from) {
  const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f);
  for (int64_t s = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; s < S;
       s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    __nv_bfloat16* r = tok + s * ld_tok;
    if (ones_channel >= 0)
      for (int i = 0; i < 4; ++i) r[ones_channel * 4 + i] = one;
    for (int i = zero_from; i < ld_tok; ++i) r[i] = zero;
  }
}

// ===================================================================== unpatchify + CFG + Euler
__global__ void __launch_bounds__(256)
unpatchify_euler_kernel(const __nv_bfloat16* __restrict__ yc, const __nv_bfloat16* __restrict__ yu, int64_t ld_y,
                        float guidance, const float* __restrict__ sigma, const float* __restrict__ sigma_next,
                        const __nv_bfloat16* __restrict__ x_t, __nv_bfloat16* __restrict__ x_next,
                        __nv_bfloat16* __restrict__ F_out, int C, int T, int H, int W) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int64_t total = static_cast<int64_t>(C) * T * Hp * Wp;
  const float sg = sigma ? *sigma : 1.0f, sn = sigma_next ? *sigma_next : 0.0f;
  const float sd = 0.5f;
  const float c_skip = (sd * sd) / (sg * sg + sd * sd);
  const float c_out = (sg * sd) / sqrtf(sg * sg + sd * sd);
  const float dt = sn - sg;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(idx % Wp);
    const int h = static_cast<int>((idx / Wp) % Hp);
    const int t = static_cast<int>((idx / (static_cast<int64_t>(Wp) * Hp)) % T);
    const int c = static_cast<int>(idx / (static_cast<int64_t>(Wp) * Hp * T));
    const int64_t s = (static_cast<int64_t>(t) * Hp + h) * Wp + w;
    const int64_t pix = ((static_cast<int64_t>(c) * T + t) * H + 2 * h) * W + 2 * w;
    float f[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // q = ph * 2 + pw; feature index q * C + c (C fastest, CleanGeneralDIT.py:709-716)
      float fc = __bfloat162float(yc[s * ld_y + q * C + c]);
      if (yu != nullptr) {
        const float fu = __bfloat162float(yu[s * ld_y + q * C + c]);
        fc = bf16_round(fc + bf16_round(guidance * bf16_round(fc - fu)));
      }
      f[q] = fc;
    }
    if (x_t != nullptr) {
      const uint32_t x0 = *reinterpret_cast<const uint32_t*>(x_t + pix);
      const uint32_t x1 = *reinterpret_cast<const uint32_t*>(x_t + pix + W);
      const float xv[4] = {bf16_lo(x0), bf16_hi(x0), bf16_lo(x1), bf16_hi(x1)};
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float den = c_skip * xv[q] + c_out * f[q];
        o[q] = xv[q] + (xv[q] - den) / sg * dt;
      }
      *reinterpret_cast<uint32_t*>(x_next + pix) = pack_bf16x2(o[0], o[1]);
      *reinterpret_cast<uint32_t*>(x_next + pix + W) = pack_bf16x2(o[2], o[3]);
    }
    if (F_out != nullptr) {
      *reinterpret_cast<uint32_t*>(F_out + pix) = pack_bf16x2(f[0], f[1]);
      *reinterpret_cast<uint32_t*>(F_out + pix + W) = pack_bf16x2(f[2], f[3]);
    }
  }
}

// ===================================================================== decode post-process -> uint8 BTHWC
__global__ void __launch_bounds__(256)
postprocess_kernel(const __nv_bfloat16* __restrict__ video, uint8_t* __restrict__ out, int64_t n_pix, int normalize_normal) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_pix;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __bfloat162float(video[c * n_pix + i]);
    uint8_t o[3];
    postprocess_pixel(v, normalize_normal, o);
#pragma unroll
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = o[c];
  }
}

// ===================================================================== stand-alone EDM scheduler ops (public scheduler API)
// mode 0: out = bf16(x * 1/sqrt(sigma^2 + 0.25))                 (model_diffusion_renderer.py:30-44)
// mode 1: out = bf16(x + (x - (c_skip x + c_out F)) / sigma * (sigma_next - sigma))   (:46-82)
__global__ void __launch_bounds__(256)
edm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ F, const float* __restrict__ sigma,
           const float* __restrict__ sigma_next, __nv_bfloat16* __restrict__ out, int64_t n, int mode) {
  const float sg = *sigma;
  const float sd = 0.5f;
  const float cin = 1.0f / sqrtf(sg * sg + sd * sd);
  const float c_skip = (sd * sd) / (sg * sg + sd * sd);
  const float c_out = (sg * sd) / sqrtf(sg * sg + sd * sd);
  const float dt = mode == 1 ? *sigma_next - sg : 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float xv = __bfloat162float(x[i]);
    float o;
    if (mode == 0) {
      o = xv * cin;
    } else {
      const float den = c_skip * xv + c_out * __bfloat162float(F[i]);
      o = xv + (xv - den) / sg * dt;
    }
    out[i] = __float2bfloat16_rn(o);
  }
}

inline int grid_for(int64_t n, int block, int max_blocks) {
  int64_t g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace
}  // namespace drb

using namespace drb;

extern "C" int drb_adaln_modulate(void* x, void* out, const void* shift, const void* scale, const void* add_gate,
                                  const void* add_vec, int rows, int D, void* stream) {
  DRB_REQUIRE(x && out && shift && scale, "null pointer");
  DRB_REQUIRE(rows > 0 && D > 0 && D % 256 == 0 && D <= 256 * kAdaMaxChunks, "D must be a multiple of 256, at most 4096");
  DRB_REQUIRE((add_gate == nullptr) == (add_vec == nullptr), "add_gate and add_vec go together");
  adaln_kernel<<<(rows + kAdaWarps - 1) / kAdaWarps, kAdaWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(shift),
      static_cast<const __nv_bfloat16*>(scale), static_cast<const __nv_bfloat16*>(add_gate),
      static_cast<const __nv_bfloat16*>(add_vec), rows, D);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_qk_norm_rope(void* qkv, int64_t ld, const void* wq, const void* wk, const void* cos_tab,
                                const void* sin_tab, int S, int num_heads, void* stream) {
  DRB_REQUIRE(qkv && wq && wk && cos_tab && sin_tab, "null pointer");
  DRB_REQUIRE(S > 0 && num_heads > 0, "S and num_heads must be positive");
  DRB_REQUIRE(ld % 4 == 0 && ld >= 3LL * num_heads * 128, "qkv pitch must hold [q | k | v] with 128-wide heads");
  qk_norm_rope_kernel<<<S, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(qkv), ld, static_cast<const __nv_bfloat16*>(wq), static_cast<const __nv_bfloat16*>(wk),
      static_cast<const __nv_bfloat16*>(cos_tab), static_cast<const __nv_bfloat16*>(sin_tab), S, num_heads);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

// One warp per layer: bound[l] = sqrt(128) * max|wq[l]| * max|wk[l]| * 1.02 (include/drb200.h)
namespace drb {
namespace {
__global__ void __launch_bounds__(32)
qk_logit_bound_kernel(const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk, float* __restrict__ bound) {
  const int l = blockIdx.x, lane = threadIdx.x;
  float mq = 0.f, mk = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mq = fmaxf(mq, fabsf(__bfloat162float(wq[l * 128 + lane * 4 + i])));
    mk = fmaxf(mk, fabsf(__bfloat162float(wk[l * 128 + lane * 4 + i])));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, o));
    mk = fmaxf(mk, __shfl_xor_sync(0xffffffffu, mk, o));
  }
  // NaN weights must not certify anything: fmaxf drops NaNs, so test them explicitly
  float nan_seen = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = __bfloat162float(wq[l * 128 + lane * 4 + i]), b = __bfloat162float(wk[l * 128 + lane * 4 + i]);
    if (a != a || b != b) nan_seen = 1.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nan_seen = fmaxf(nan_seen, __shfl_xor_sync(0xffffffffu, nan_seen, o));
  if (lane == 0) bound[l] = nan_seen > 0.f ? INFINITY : 11.313708498984761f * mq * mk * 1.02f;
}
}  // namespace
}  // namespace drb

extern "C" int drb_qk_logit_bound(const void* wq, const void* wk, float* bound, int layers, void* stream) {
  using namespace drb;
  DRB_REQUIRE(wq && wk && bound && layers > 0, "bad arguments");
  qk_logit_bound_kernel<<<layers, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(wq),
                                                                            static_cast<const __nv_bfloat16*>(wk), bound);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_gemv_bf16_batched(const void* W, int64_t ldw, int64_t w_batch_stride, const void* x,
                                     int64_t x_batch_stride, void* y, int64_t y_batch_stride, const void* add,
                                     int64_t add_batch_stride, int count, int N, int K, int act, void* stream) {
  DRB_REQUIRE(W && x && y, "null pointer");
  DRB_REQUIRE(count > 0 && N > 0 && K > 0, "count, N, K must be positive");
  DRB_REQUIRE(K % 8 == 0 && K <= kGemvMaxK, "K must be a multiple of 8, at most 4096");
  DRB_REQUIRE(ldw % 8 == 0 && ldw >= K && w_batch_stride % 8 == 0, "weight pitch / batch stride must be multiples of 8");
  DRB_REQUIRE(act == 0 || act == 1, "act must be 0 (identity) or 1 (SiLU)");
  dim3 grid((N + kGemvWarps - 1) / kGemvWarps, count);
  gemv_kernel<<<grid, kGemvWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(W), ldw, w_batch_stride, static_cast<const __nv_bfloat16*>(x), x_batch_stride,
      static_cast<__nv_bfloat16*>(y), y_batch_stride, static_cast<const __nv_bfloat16*>(add), add_batch_stride, N, K, act);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_gemv_bf16(const void* W, int64_t ldw, const void* x, void* y, const void* add, int N, int K, int act,
                             void* stream) {
  return drb_gemv_bf16_batched(W, ldw, 0, x, 0, y, 0, add, 0, 1, N, K, act, stream);
}

extern "C" int drb_sigma_embedding(const float* sigma, const void* w_aff, void* e_out, void* emb_out, int D, void* stream) {
  DRB_REQUIRE(sigma && w_aff && e_out && emb_out, "null pointer");
  DRB_REQUIRE(D > 0 && D % 2 == 0, "D must be even");
  sigma_embedding_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      sigma, static_cast<const __nv_bfloat16*>(w_aff), static_cast<__nv_bfloat16*>(e_out),
      static_cast<__nv_bfloat16*>(emb_out), D);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_scale_patchify(const void* x_t, const float* sigma, void* tokens, int64_t ld_tok, int C, int T, int H,
                                  int W, void* stream) {
  DRB_REQUIRE(x_t && sigma && tokens, "null pointer");
  DRB_REQUIRE(C > 0 && T > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad latent shape");
  DRB_REQUIRE(ld_tok % 4 == 0 && ld_tok >= 4 * C, "token pitch too small");
  const int64_t total = static_cast<int64_t>(C) * T * (H / 2) * (W / 2);
  patchify_kernel<true><<<grid_for(total, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x_t), sigma, static_cast<__nv_bfloat16*>(tokens), ld_tok, 0, C, T, H, W);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_patchify_condition(const void* src, void* tokens, int64_t ld_tok, int c0, int C, int T, int H, int W,
                                      int ones_channel, int zero_from, void* stream) {
  DRB_REQUIRE(tokens, "null pointer");
  DRB_REQUIRE(C >= 0 && T > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad latent shape");
  DRB_REQUIRE(ld_tok % 4 == 0 && ld_tok >= 4 * (c0 + C), "token pitch too small");
  DRB_REQUIRE(ones_channel < 0 || 4 * (ones_channel + 1) <= ld_tok, "ones_channel outside the token row");
  DRB_REQUIRE(zero_from >= 0 && zero_from <= ld_tok, "zero_from outside the token row");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t S = static_cast<int64_t>(T) * (H / 2) * (W / 2);
  if (C > 0) {
    DRB_REQUIRE(src, "null pointer");
    patchify_kernel<false><<<grid_for(S * C, 256, 148 * 16), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(src), nullptr, static_cast<__nv_bfloat16*>(tokens), ld_tok, c0, C, T, H, W);
    DRB_CUDA(cudaGetLastError());
  }
  if (ones_channel >= 0 || zero_from < ld_tok) {
    patchify_fill_kernel<<<grid_for(S, 256, 148 * 16), 256, 0, s>>>(static_cast<__nv_bfloat16*>(tokens), ld_tok, S,
                                                                     ones_channel, zero_from);
    DRB_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int drb_unpatchify_euler(const void* y_cond, const void* y_uncond, int64_t ld_y, float guidance,
                                    const float* sigma, const float* sigma_next, const void* x_t, void* x_next,
                                    void* F_out, int C, int T, int H, int W, void* stream) {
  DRB_REQUIRE(y_cond, "null pointer");
  DRB_REQUIRE((x_t == nullptr) == (x_next == nullptr), "x_t and x_next go together");
  DRB_REQUIRE(x_t != nullptr || F_out != nullptr, "nothing to write: pass x_t/x_next and/or F_out");
  DRB_REQUIRE(x_t == nullptr || (sigma && sigma_next), "the Euler update needs sigma and sigma_next");
  DRB_REQUIRE(C > 0 && T > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad latent shape");
  DRB_REQUIRE(ld_y >= 4 * C, "y pitch too small");
  const int64_t total = static_cast<int64_t>(C) * T * (H / 2) * (W / 2);
  unpatchify_euler_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y_cond), static_cast<const __nv_bfloat16*>(y_uncond), ld_y, guidance, sigma,
      sigma_next, static_cast<const __nv_bfloat16*>(x_t), static_cast<__nv_bfloat16*>(x_next),
      static_cast<__nv_bfloat16*>(F_out), C, T, H, W);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_postprocess_u8(const void* video, void* out_u8, int T, int H, int W, int normalize_normal, void* stream) {
  DRB_REQUIRE(video && out_u8, "null pointer");
  DRB_REQUIRE(T > 0 && H > 0 && W > 0, "bad video shape");
  const int64_t n_pix = static_cast<int64_t>(T) * H * W;
  postprocess_kernel<<<grid_for(n_pix, 256, 148 * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(video), static_cast<uint8_t*>(out_u8), n_pix, normalize_normal);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_edm_scale_input(const void* x, const float* sigma, void* out, int64_t n, void* stream) {
  DRB_REQUIRE(x && sigma && out && n > 0, "null pointer or empty tensor");
  edm_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), nullptr, sigma, nullptr, static_cast<__nv_bfloat16*>(out), n, 0);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_edm_euler_step(const void* model_output, const void* x, const float* sigma, const float* sigma_next,
                                  void* out, int64_t n, void* stream) {
  DRB_REQUIRE(model_output && x && sigma && sigma_next && out && n > 0, "null pointer or empty tensor");
  edm_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(model_output), sigma, sigma_next,
      static_cast<__nv_bfloat16*>(out), n, 1);
  DRB_CUDA(cudaGetLastError());
  return 0;
}
