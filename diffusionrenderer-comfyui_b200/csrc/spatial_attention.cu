// drb_spatial_attention_d512 — the mid-block spatial attention of the CV8x8x8 tokenizer (CosmosCausalAttention over the H*W
// tokens of one latent frame, ONE head of dim 512; reference CleanVAE.py:50-51,59-60 -> diffusers AutoencoderKLCosmos) as one
// flash-attention kernel on tcgen05 tensor cores, instead of scores GEMM -> row softmax -> transpose -> P.V GEMM with a
// 396 MB score matrix per frame written once and read twice.
//
// A head of 512 does not fit the usual flash tile: O alone (128 rows x 512 fp32) would fill all 512 TMEM columns.  So one
// CTA owns 128 query rows and ONE HALF (256 columns) of V / O; the other half is a second CTA (blockIdx.y) that recomputes
// the same scores.  Executed MMA work is 1.5x the algorithmic work (QK^T twice, P.V once) — the price of the head width —
// against 5 HBM / L2 passes over a 396 MB score matrix before.
//
//   warp 0      TMA producer: Q once (8 chunks of 128 rows x 64 dims = 128 KB), then per key tile of 128 keys the 8 K chunks
//               and the 4 V chunks (128 keys x 64 columns) through a ring of 16 KB slots
//   warp 1      tcgen05.mma issuer: S_b = Q K_j^T (32 k-steps, N = 128) into one of TWO score buffers, so that Q K_{j+1}^T runs
//               while the softmax warps work on S_j;  O[:, c] += P_j V_j[:, c] for the four 64-column chunks (A = P_j in TMEM)
//   warp 2      TMEM allocator: S_0 | S_1 | O (128 + 128 + 256 columns)
//   warps 4..7  softmax (one query row per thread): per-tile row maximum, lazy reference (moves only when a logit exceeds it
//               by more than 2^8), P written over S as packed bf16, O rescaled in TMEM when the reference moves
// Roofline: tensor pipe; algorithmic work 4 * n^2 * 512 flop per frame.
#include <math.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace drb {
namespace {

constexpr int kD = 512;               // head dim (q.k reduction length and V width)
constexpr int kVHalf = 256;           // V / O columns per CTA
constexpr int kTile = 128;            // query rows per CTA, keys per tile
constexpr int kChunkBytes = kTile * 64 * 2;       // 16 KB: 128 rows x 64 bf16 (one SWIZZLE_128B box)
constexpr int kQChunks = kD / 64;     // 8
constexpr int kVChunks = kVHalf / 64; // 4
constexpr int kSlots = 5;
constexpr int kThreads = 256;
constexpr int kSmem = kQChunks * kChunkBytes + kSlots * kChunkBytes + 1024 + 256;
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
constexpr uint32_t kLoK = 1u << 16;                                    // K-major: LBO field unused
constexpr uint32_t kLoMn = (kChunkBytes >> 4) << 16;                   // MN-major: LBO between 64-wide chunks (one chunk only here)

struct SpatialParams {
  __nv_bfloat16* out;
  int64_t ld_o;
  int n;            // tokens per frame
  float scale_log2; // log2(e) / sqrt(512)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }

__global__ void __launch_bounds__(kThreads, 1)
spatial_attention_kernel(const __grid_constant__ CUtensorMap tmap, const SpatialParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_ring = smem + kQChunks * kChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ring + kSlots * kChunkBytes);
  uint64_t* q_full = bars;
  uint64_t* ring_full = bars + 1;
  uint64_t* ring_empty = ring_full + kSlots;
  uint64_t* s_full = ring_empty + kSlots;   // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* pv_done = p_full + 2;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTile;
  const int half = blockIdx.y;
  const int frame = blockIdx.z;
  const int n_kv = (p.n + kTile - 1) / kTile;

  if (warp_idx == 0 && lane == 0) prefetch_tmap(&tmap);
  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);   // one arrival per softmax warp
    }
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQChunks * kChunkBytes);
      for (int c = 0; c < kQChunks; ++c) tma_load_3d(smem_q + c * kChunkBytes, &tmap, q_full, c * 64, q0, frame);
      int slot = 0;
      uint32_t phase = 0;
      auto load = [&](int col, int row) {
        mbar_wait(&ring_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&ring_full[slot], kChunkBytes);
        tma_load_3d(smem_ring + slot * kChunkBytes, &tmap, &ring_full[slot], col, row, frame);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      };
      for (int c = 0; c < kQChunks; ++c) load(kD + c * 64, 0);                              // K_0
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv)
          for (int c = 0; c < kQChunks; ++c) load(kD + c * 64, (j + 1) * kTile);            // K_{j+1}
        for (int c = 0; c < kVChunks; ++c) load(2 * kD + half * kVHalf + c * 64, j * kTile); // V_j (this CTA's half)
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp in uniform control flow, one lane issues)
    constexpr uint32_t idesc_qk = make_idesc_bf16(kTile, kTile, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, 64, false, true);   // B = V chunk, MN-major
    const bool leader = elect_one();
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t q_lo = ((smem_u32(smem_q) & 0x3FFFF) >> 4) | kLoK;
    const uint32_t ring_k = ((smem_u32(smem_ring) & 0x3FFFF) >> 4) | kLoK;
    const uint32_t ring_v = ((smem_u32(smem_ring) & 0x3FFFF) >> 4) | kLoMn;
    int slot = 0;
    uint32_t phase = 0;
    auto qk = [&](int sb) {      // S_sb = Q K^T over the 8 chunks that come next in the ring
      for (int c = 0; c < kQChunks; ++c) {
        mbar_wait(&ring_full[slot], phase);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss<1>(tm + sb * kTile, desc64(q_lo + c * (kChunkBytes >> 4) + kk * 2),
                       desc64(ring_k + slot * (kChunkBytes >> 4) + kk * 2), idesc_qk, (c | kk) != 0);
          umma_commit(&ring_empty[slot]);
        }
        __syncwarp();
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&s_full[sb]);
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    qk(0);
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      if (j + 1 < n_kv) qk(sb ^ 1);                       // runs while the softmax warps work on S_sb
      mbar_wait(&p_full[sb], (j >> 1) & 1);
      tc_fence_after();
      for (int c = 0; c < kVChunks; ++c) {                // O[:, 64c .. 64c+63] (+)= P V_j[:, chunk c]
        mbar_wait(&ring_full[slot], phase);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k)
            umma_ts(tm + 2 * kTile + c * 64, tm + sb * kTile + k * 8, desc64(ring_v + slot * (kChunkBytes >> 4) + k * (2048 >> 4)),
                    idesc_pv, (j | k) != 0);
          umma_commit(&ring_empty[slot]);
        }
        __syncwarp();
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
      if (leader) umma_commit(pv_done);                   // O holds the contributions of tiles 0..j
      __syncwarp();
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------------ softmax warps
    const int quad = warp_idx & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + 2 * kTile;
    float M = 0.f, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      const uint32_t s_tmem = tmem_base + lane_off + sb * kTile;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld32(s_tmem, s0);
      tmem_ld32(s_tmem + 32, s1);
      tmem_ld32(s_tmem + 64, s2);
      tmem_ld32(s_tmem + 96, s3);
      tmem_wait_ld();
      const int valid = p.n - j * kTile;
      if (valid < kTile) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) s0[i] = 0xff800000u;
          if (32 + i >= valid) s1[i] = 0xff800000u;
          if (64 + i >= valid) s2[i] = 0xff800000u;
          if (96 + i >= valid) s3[i] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 1])));
        mx = fmaxf(mx, fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1])));
        mx = fmaxf(mx, fmaxf(__uint_as_float(s2[i]), __uint_as_float(s2[i + 1])));
        mx = fmaxf(mx, fmaxf(__uint_as_float(s3[i]), __uint_as_float(s3[i + 1])));
      }
      const float m_tile = mx * p.scale_log2;
      if (j == 0) {
        M = m_tile;
      } else {
        const bool need = m_tile - M > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O must be quiescent: P V_{j-1} has completed (its commit is phase j-1 of pv_done)
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
          const float alpha = need ? ex2f(M - m_tile) : 1.0f;
          if (need) M = m_tile;
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < kVHalf / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(o_tmem + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(o_tmem + c * 32, o);
          }
          tmem_wait_st();
        }
      }
      const float negm = -M;
      float sum = 0.f;
      uint32_t pk[32], pk2[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = ex2f(fmaf(__uint_as_float(s0[2 * i]), p.scale_log2, negm)), b = ex2f(fmaf(__uint_as_float(s0[2 * i + 1]), p.scale_log2, negm));
        const float c = ex2f(fmaf(__uint_as_float(s1[2 * i]), p.scale_log2, negm)), d = ex2f(fmaf(__uint_as_float(s1[2 * i + 1]), p.scale_log2, negm));
        sum += (a + b) + (c + d);
        pk[i] = pack_bf16x2(a, b);
        pk[16 + i] = pack_bf16x2(c, d);
      }
      tmem_st32(s_tmem, pk);             // P columns [0,32): keys 0..63
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = ex2f(fmaf(__uint_as_float(s2[2 * i]), p.scale_log2, negm)), b = ex2f(fmaf(__uint_as_float(s2[2 * i + 1]), p.scale_log2, negm));
        const float c = ex2f(fmaf(__uint_as_float(s3[2 * i]), p.scale_log2, negm)), d = ex2f(fmaf(__uint_as_float(s3[2 * i + 1]), p.scale_log2, negm));
        sum += (a + b) + (c + d);
        pk2[i] = pack_bf16x2(a, b);
        pk2[16 + i] = pack_bf16x2(c, d);
      }
      tmem_st32(s_tmem + 32, pk2);       // P columns [32,64): keys 64..127
      l += sum;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }
    // ---- epilogue: O / l -> bf16 -> out[frame*n + row, half*256 ...]
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const int row = q0 + quad * 32 + lane;
    const float inv_l = 1.0f / l;
    __nv_bfloat16* dst = p.out + (static_cast<int64_t>(frame) * p.n + (row < p.n ? row : 0)) * p.ld_o + half * kVHalf;
#pragma unroll 1
    for (int c = 0; c < kVHalf / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c * 32, o);
      tmem_wait_ld();
      if (row < p.n) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack_bf16x2(__uint_as_float(o[g * 8 + 2 * i]) * inv_l, __uint_as_float(o[g * 8 + 2 * i + 1]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace
}  // namespace drb

extern "C" int drb_spatial_attention_d512(const void* qkv, int64_t ld, void* out, int64_t ld_o, int frames, int n, void* stream) {
  using namespace drb;
  DRB_REQUIRE(qkv && out, "null pointer");
  DRB_REQUIRE(frames > 0 && frames <= 65535 && n > 0, "bad sizes");
  DRB_REQUIRE(ld % 8 == 0 && ld >= 3 * kD && ld_o % 8 == 0 && ld_o >= kD, "row pitches must be multiples of 8 and hold q | k | v (3 x 512) / 512 columns");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out not 16-byte aligned");
  // one 3-D map over [frames][n][3*512]: a box never crosses a frame, rows past the frame's last token read as zeros
  CUtensorMap tm;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * kD), static_cast<uint64_t>(n), static_cast<uint64_t>(frames)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * static_cast<uint64_t>(n)};
  const uint32_t box[3] = {64, kTile, 1};
  int rc = make_tmap_nd_bf16(&tm, qkv, 3, dims, strides, box);
  if (rc) return rc;
  static DeviceOnce configured;
  rc = device_once(configured, [] {
    return check_cuda(cudaFuncSetAttribute(spatial_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem), "spatial_attention smem");
  });
  if (rc) return rc;
  SpatialParams p{};
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_o = ld_o;
  p.n = n;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(kD));
  dim3 grid((n + kTile - 1) / kTile, 2, frames);
  spatial_attention_kernel<<<grid, kThreads, kSmem, static_cast<cudaStream_t>(stream)>>>(tm, p);
  DRB_CUDA(cudaGetLastError());
  return 0;
}
