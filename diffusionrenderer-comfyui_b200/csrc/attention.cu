// drb_attention_bf16 — non-causal softmax(Q K^T / sqrt(128)) V for head_dim 128 on tcgen05 tensor cores.
//
// Replaces PytorchDotProductAttention.forward / F.scaled_dot_product_attention (CleanGeneralDIT.py:181-203) and
// produces the head-flattened (S, H*128) layout that to_out consumes (SURVEY.md defect D1).
//
// One CTA = one head x 256 query rows (two 128-row tiles that ping-pong on the tensor pipe):
//   warp 0        TMA producer: Q once, then K_0 V_0 K_1 V_1 ... through a 5-slot ring of 32 KB tiles
//   warp 1        tcgen05.mma issuer (one thread):  S_t = Q_t K_j^T (SS),  O_t += P_t V_j (A = P_t in TMEM, B = V_j MN-major)
//   warp 2        TMEM allocator (all 512 columns: S_0 | S_1 | O_0 | O_1, each 128 fp32 columns; P_t aliases S_t)
//   warps 4..7    softmax warpgroup of tile 0 (one query row per thread)
//   warps 8..11   softmax warpgroup of tile 1
// While warpgroup t exponentiates S_t(j), the tensor pipe runs P V and Q K^T of the other tile.  The running max is
// lazy: O_t (in TMEM) is only rescaled when a row maximum grows by more than 2^8, which the owning softmax thread
// does itself between s_full and p_full, when O_t is quiescent.
//
// Roofline: tensor pipe, 4*q_len*kv_len*128 flop per head.
#include "../../include/drb200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace drb {
namespace {

constexpr int kHeadDim = 128;
constexpr int kTileQ = 128;
constexpr int kTileKV = 128;
constexpr int kQTilesPerCta = 2;
constexpr int kTileBytes = kTileKV * kHeadDim * 2;   // 32 KB: two 64-column TMA boxes of 16 KB
constexpr int kBoxBytes = kTileBytes / 2;
constexpr int kKvSlots = 5;
constexpr int kAttnThreads = 384;
constexpr int kAttnSmem = kQTilesPerCta * kTileBytes + kKvSlots * kTileBytes + 1024 + 256;
constexpr float kScaleLog2 = 0.08838834764831845f * 1.4426950408889634f;   // log2(e) / sqrt(128)
constexpr float kRescaleThreshold = 8.0f;   // in log2 units: P stays below 2^8

struct AttnParams {
  __nv_bfloat16* o;
  int64_t ld_o;
  int q_len, kv_len;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kRegs>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// UMMA shared-memory descriptors are built as {hi: constant per layout, lo: (address >> 4) | LBO field}, so the issue
// loop only does 32-bit adds.  K-major 128 x 128 bf16 tile = two 128-row x 128-byte swizzled boxes; k-step `k` covers
// 16 columns: box k / 4, byte offset (k % 4) * 32 inside the swizzle atom.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
constexpr uint32_t kLoKMajor = 1u << 16;                               // LBO field (ignored for swizzled K-major)
constexpr uint32_t kLoMnMajor = (kBoxBytes >> 4) << 16;                // LBO = 16 KB between the two 64-wide MN chunks
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }
__device__ __forceinline__ constexpr uint32_t kstep_off(int k) { return ((k >> 2) * kBoxBytes + (k & 3) * 32) >> 4; }

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + kQTilesPerCta * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + kKvSlots * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kKvSlots;
  uint64_t* s_full = kv_empty + kKvSlots;   // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_full + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (kTileQ * kQTilesPerCta);
  const int n_kv = (p.kv_len + kTileKV - 1) / kTileKV;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvSlots; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);   // one arrival per softmax warp
    }
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

  if (warp_idx < 4) {
    reg_dec<56>();
    if (warp_idx == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      const int col = head * kHeadDim;
      mbar_arrive_expect_tx(q_full, kQTilesPerCta * kTileBytes);
      for (int t = 0; t < kQTilesPerCta; ++t)
        for (int b = 0; b < 2; ++b)
          tma_load_2d(smem_q + t * kTileBytes + b * kBoxBytes, &tmap_q, q_full, col + b * 64, q0 + t * kTileQ);
      int slot = 0;
      uint32_t phase = 0;
      for (int i = 0; i < 2 * n_kv; ++i) {
        mbar_wait(&kv_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[slot], kTileBytes);
        const CUtensorMap* tm = (i & 1) ? &tmap_v : &tmap_k;
        uint8_t* dst = smem_kv + slot * kTileBytes;
        const int row = (i >> 1) * kTileKV;
        tma_load_2d(dst, tm, &kv_full[slot], col, row);
        tma_load_2d(dst + kBoxBytes, tm, &kv_full[slot], col + 64, row);
        if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
      }
    } else if (warp_idx == 1 && lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_qk = make_idesc_bf16(kTileQ, kTileKV, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, kHeadDim, false, true);   // B = V is MN-major
      uint32_t q_lo = ((smem_u32(smem_q) & 0x3FFFF) >> 4) | kLoKMajor;
      const uint32_t kv_lo = ((smem_u32(smem_kv) & 0x3FFFF) >> 4) | kLoKMajor;
      const uint32_t v_lo = ((smem_u32(smem_kv) & 0x3FFFF) >> 4) | kLoMnMajor;
      int slot = 0;
      uint32_t phase = 0;
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[slot], phase);
      tc_fence_after();
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss<1>(tmem_base + t * 128, desc64(q_lo + t * (kTileBytes >> 4) + kstep_off(k)),
                     desc64(kv_lo + slot * (kTileBytes >> 4) + kstep_off(k)), idesc_qk, k != 0);
        umma_commit(&s_full[t]);
      }
      umma_commit(&kv_empty[slot]);
      if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
      for (int j = 0; j < n_kv; ++j) {
        const int v_slot = slot;
        mbar_wait(&kv_full[v_slot], phase);
        if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
        const bool more = j + 1 < n_kv;
        asm volatile("" : "+r"(q_lo));   // keep the 16 Q descriptors out of (spilled) loop-invariant registers
        int k_slot = 0;
        if (more) {
          k_slot = slot;
          mbar_wait(&kv_full[k_slot], phase);
          if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
        }
        tc_fence_after();
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], j & 1);
          tc_fence_after();
          // O_t (+)= P_t V_j : A = P_t (bf16 pairs packed in TMEM columns, 8 columns per k-step of 16),
          //                    B = 16 rows of V (2048 B) x 128 columns (two 64-wide chunks 16 KB apart)
#pragma unroll
          for (int k = 0; k < kTileKV / 16; ++k)
            umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + k * 8,
                    desc64(v_lo + v_slot * (kTileBytes >> 4) + k * (2048 >> 4)), idesc_pv, (j | k) != 0);
          if (t == 1) umma_commit(&kv_empty[v_slot]);
          if (more) {
#pragma unroll
            for (int k = 0; k < kHeadDim / 16; ++k)
              umma_ss<1>(tmem_base + t * 128, desc64(q_lo + t * (kTileBytes >> 4) + kstep_off(k)),
                         desc64(kv_lo + k_slot * (kTileBytes >> 4) + kstep_off(k)), idesc_qk, k != 0);
          }
          umma_commit(&s_full[t]);   // S_t(j+1) ready (and P_t V_j done); after the last tile: O_t final
          if (more && t == 1) umma_commit(&kv_empty[k_slot]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    reg_inc<224>();
    const int t = (warp_idx - 4) >> 2;             // query tile of this warpgroup
    const int quad = warp_idx & 3;                 // TMEM lane quadrant
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + t * 128;
    const uint32_t o_tmem = tmem_base + lane_off + 256 + t * 128;
    float m_ref = 0.f, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld32(s_tmem, s0);
      tmem_ld32(s_tmem + 32, s1);
      tmem_ld32(s_tmem + 64, s2);
      tmem_ld32(s_tmem + 96, s3);
      tmem_wait_ld();
      const int valid = p.kv_len - j * kTileKV;    // >= 128 except on a ragged last tile
      if (valid < kTileKV) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) s0[i] = 0xff800000u;
          if (32 + i >= valid) s1[i] = 0xff800000u;
          if (64 + i >= valid) s2[i] = 0xff800000u;
          if (96 + i >= valid) s3[i] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(s0[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
        mx2 = fmaxf(mx2, __uint_as_float(s2[i]));
        mx3 = fmaxf(mx3, __uint_as_float(s3[i]));
      }
      const float m_cur = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (j == 0) {
        m_ref = m_cur;
      } else {
        const float m_new = fmaxf(m_ref, m_cur);
        const bool grow = (m_new - m_ref) * kScaleLog2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {
          // rescale the running sum and this row of O_t (quiescent: P_t V_{j-1} completed before s_full fired)
          const float alpha = ex2((m_ref - m_new) * kScaleLog2);
          l *= alpha;
          m_ref = m_new;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
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
      const float neg_m = -m_ref * kScaleLog2;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = ex2(fmaf(__uint_as_float(s0[2 * i]), kScaleLog2, neg_m));
        const float b = ex2(fmaf(__uint_as_float(s0[2 * i + 1]), kScaleLog2, neg_m));
        const float c = ex2(fmaf(__uint_as_float(s1[2 * i]), kScaleLog2, neg_m));
        const float d = ex2(fmaf(__uint_as_float(s1[2 * i + 1]), kScaleLog2, neg_m));
        sum0 += a; sum1 += b; sum2 += c; sum3 += d;
        pk[i] = pack_bf16x2(a, b);
        pk[16 + i] = pack_bf16x2(c, d);
      }
      tmem_st32(s_tmem, pk);        // P_t columns [0,32): keys 0..63
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = ex2(fmaf(__uint_as_float(s2[2 * i]), kScaleLog2, neg_m));
        const float b = ex2(fmaf(__uint_as_float(s2[2 * i + 1]), kScaleLog2, neg_m));
        const float c = ex2(fmaf(__uint_as_float(s3[2 * i]), kScaleLog2, neg_m));
        const float d = ex2(fmaf(__uint_as_float(s3[2 * i + 1]), kScaleLog2, neg_m));
        sum0 += a; sum1 += b; sum2 += c; sum3 += d;
        pk[i] = pack_bf16x2(a, b);
        pk[16 + i] = pack_bf16x2(c, d);
      }
      tmem_st32(s_tmem + 32, pk);   // P_t columns [32,64): keys 64..127
      l += (sum0 + sum1) + (sum2 + sum3);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // ---- epilogue: O_t / l -> bf16 -> global (row = one thread, 256 contiguous bytes per head)
    mbar_wait(&s_full[t], n_kv & 1);
    tc_fence_after();
    const int row = q0 + t * kTileQ + quad * 32 + lane;
    const float inv_l = 1.0f / l;
    __nv_bfloat16* dst = p.o + static_cast<int64_t>(row) * p.ld_o + head * kHeadDim;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c * 32, o);
      tmem_wait_ld();
      if (row < p.q_len) {
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

extern "C" int drb_attention_bf16(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                                  int q_len, int kv_len, int num_heads, void* stream) {
  using namespace drb;
  DRB_REQUIRE(q && k && v && o, "null pointer");
  DRB_REQUIRE(q_len > 0 && kv_len > 0 && num_heads > 0, "q_len, kv_len, num_heads must be positive");
  DRB_REQUIRE(ld_qkv % 8 == 0 && ld_o % 8 == 0, "row pitches must be multiples of 8 elements");
  DRB_REQUIRE(ld_qkv >= static_cast<int64_t>(num_heads) * kHeadDim && ld_o >= static_cast<int64_t>(num_heads) * kHeadDim,
              "row pitch smaller than num_heads * 128");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0, "o not 16-byte aligned");
  CUtensorMap tq, tk, tv;
  const uint64_t cols = static_cast<uint64_t>(num_heads) * kHeadDim;
  int rc = make_tmap_2d_bf16(&tq, q, q_len, cols, ld_qkv, kTileQ, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tk, k, kv_len, cols, ld_qkv, kTileKV, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tv, v, kv_len, cols, ld_qkv, kTileKV, 64);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    DRB_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured = true;
  }
  AttnParams p{static_cast<__nv_bfloat16*>(o), ld_o, q_len, kv_len};
  dim3 grid((q_len + kTileQ * kQTilesPerCta - 1) / (kTileQ * kQTilesPerCta), num_heads);
  attention_kernel<<<grid, kAttnThreads, kAttnSmem, static_cast<cudaStream_t>(stream)>>>(tq, tk, tv, p);
  DRB_CUDA(cudaGetLastError());
  return 0;
}
