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
// While warpgroup t exponentiates S_t(j), the tensor pipe runs P V and Q K^T of the other tile.  Two softmax flavours,
// chosen per launch by a uniform branch on a device-resident certificate (so no host synchronisation is needed to pick):
//   certified  the reference is lazy and max-free: tile 0 fixes it at the exact row maximum, later it moves (and O_t, in
//              TMEM, is rescaled by the owning softmax thread between s_full and p_full, when O_t is quiescent) only after
//              a tile whose row sum reached 2^8.  Exact whenever the logits are bounded: the caller passes a device float
//              B with |q.k| / sqrt(128) <= B (the DiT computes it from its q/k RMSNorm weights, drb_qk_logit_bound); for
//              B <= kMaxCertifiedLogit, 2^(x - M) cannot overflow however the keys are ordered.
//   safe       a per-tile row maximum moves the reference whenever a logit exceeds it by more than 2^8 — the general
//              form for unbounded logits (no certificate, or B too large), a few percent slower.
//
// Roofline: tensor pipe, 4*q_len*kv_len*128 flop per head.
#include <stdlib.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "cp_sync.cuh"
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
// |q.k| / sqrt(128) bound (natural units) up to which the max-free softmax is provably overflow-free: the reference sits at
// tile 0's row maximum M0 >= -B log2(e) and only ever moves up, so x = s*c - M <= 2 B log2(e) <= 113 < 127 for B <= 39.
constexpr float kMaxCertifiedLogit = 39.0f;

struct AttnParams {
  // output row r goes to o_peers[r / rows_per_rank] at local row r % rows_per_rank, column col0 + head*128: a single
  // destination for the one-GPU path; under context parallelism (cp.cu) the GPU that owns the token (P2P store)
  void* o_peers[DRB_CP_MAX_RANKS];
  int64_t ld_o;
  int q_len, kv_len;
  int rows_per_rank, col0;
  // batched sequences (cp only): "head" index = b * heads_per_batch + h; sequence b's rows start at local row b * batch_rows
  // of the destination and its head h sits at column col0 + h*128.  Unbatched: heads_per_batch = num_heads.
  int heads_per_batch, batch_rows;
  // certificate of the max-free softmax: device float B >= |q.k| / sqrt(128) (NULL = none); force_safe: 1 always the
  // per-tile-max softmax, 0 by certificate, -1 never (tuning only)
  const float* logit_bound;
  int force_safe;
  // context parallelism: wait for the peers' q / k / v stores before the first load, signal the peers once every CTA has
  // stored its output rows (csrc/cp_sync.cuh); off unless a drb_cp_sync descriptor was passed
  CpSync sync;
  // ring form (drb_attention_bf16_ring): this launch is one K/V block of a longer key sequence.  The running state per
  // (row, head) — reference m (log2 domain), sum l, un-normalised fp32 output — lives in ring_ml [q_len, H, 2] and ring_o
  // [q_len, H*128]; the epilogue merges this block into it and, on the last block, writes the normalised bf16 rows.
  float* ring_o;
  float* ring_ml;
  int ring_first, ring_last, num_heads;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // not volatile: ptxas may batch the MUFUs
  return y;
}

// ---- packed fp32x2 arithmetic (one issue slot for two lanes of work on sm_100) and the FMA-pipe exp2
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^x for two values on the FMA pipe (the MUFU unit is the co-bottleneck of the softmax):
// n = round(x) by the 1.5*2^23 trick, f = x - n in [-0.5, 0.5], degree-3 polynomial (max rel. error 2.4e-4, an order
// of magnitude below the bf16 rounding P receives), exponent patched in with one integer shift-add.
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& r0, float& r1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x2 = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t magic = pack2(12582912.0f, 12582912.0f), neg_magic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t xr = add2(x2, magic);
  const uint64_t nf = add2(xr, neg_magic);
  const uint64_t f = fma2(nf, pack2(-1.0f, -1.0f), x2);
  uint64_t p = fma2(pack2(0.05273903161287308f, 0.05273903161287308f), f, pack2(0.24209719896316528f, 0.24209719896316528f));
  p = fma2(p, f, pack2(0.6935965418815613f, 0.6935965418815613f));
  p = fma2(p, f, pack2(0.9999658465385437f, 0.9999658465385437f));
  float p0, p1, n0, n1;
  unpack2(p, p0, p1);
  unpack2(xr, n0, n1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(n0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(n1) << 23));
}
// kPolyMask: which key pairs (index i within a 32-column chunk, 16 pairs) take the polynomial instead of MUFU.EX2
// P for one 32-column chunk of S: 16 packed bf16x2 words into pk[], partial row sums into sum2
template <uint32_t kPolyMask>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&s)[32], uint64_t scale2, uint64_t negm2, uint32_t* pk,
                                              uint64_t& sum2) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t x2 = fma2(pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), scale2, negm2);
    float a, b;
    if ((kPolyMask >> i) & 1u) {
      exp2_poly2(x2, a, b);
    } else {
      float x0, x1;
      unpack2(x2, x0, x1);
      a = ex2(x0);
      b = ex2(x1);
    }
    sum2 = add2(sum2, pack2(a, b));
    pk[i] = pack_bf16x2(a, b);
  }
}


// UMMA shared-memory descriptors are built as {hi: constant per layout, lo: (address >> 4) | LBO field}, so the issue
// loop only does 32-bit adds.  K-major 128 x 128 bf16 tile = two 128-row x 128-byte swizzled boxes; k-step `k` covers
// 16 columns: box k / 4, byte offset (k % 4) * 32 inside the swizzle atom.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
constexpr uint32_t kLoKMajor = 1u << 16;                               // LBO field (ignored for swizzled K-major)
constexpr uint32_t kLoMnMajor = (kBoxBytes >> 4) << 16;                // LBO = 16 KB between the two 64-wide MN chunks
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }
__device__ __forceinline__ constexpr uint32_t kstep_off(int k) { return ((k >> 2) * kBoxBytes + (k & 3) * 32) >> 4; }

#ifdef DRB_ATTN_PROFILE
// phase cycle counters of CTA (0,0): [0..7] softmax warp 4, [8..15] MMA thread (tuning builds only)
__device__ unsigned long long g_attn_prof[16];
#define PROF_DECL unsigned long long prof_t = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF(i) do { unsigned long long _n = clock64(); prof_acc[i] += _n - prof_t; prof_t = _n; } while (0)
#define PROF_DUMP(base) do { if (blockIdx.x == 0 && blockIdx.y == 0) for (int _i = 0; _i < 8; ++_i) g_attn_prof[(base) + _i] = prof_acc[_i]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_DUMP(base)
#endif

// kPair: the two CTAs of a cluster work on adjacent query blocks of the same head and need the same K/V tiles in the same
// order: each loads half of every tile (64 rows) and TMA multicasts it into both CTAs, halving the L2 -> SM traffic; a ring
// slot is reused only after the MMAs of BOTH CTAs have released it (multicast tcgen05.commit on a 2-arrival barrier).
template <uint32_t kPolyMask, bool kPair>
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
  uint64_t* p_full = s_full + 2;            // [2 tiles][2 halves of the key range]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_full + 4);

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
      mbar_init(&kv_empty[i], kPair ? 2 : 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[2 * t], 4);   // one arrival per softmax warp
      mbar_init(&p_full[2 * t + 1], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync(); else __syncthreads();   // the peer's barriers must exist before anything lands on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t pair_rank = kPair ? cluster_ctarank() : 0u;

  if (warp_idx < 4) {
    reg_dec<56>();
    if (warp_idx == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      const int col = head * kHeadDim;
      cp_wait(p.sync);   // q / k / v rows stored by peer GPUs (the QKV GEMM epilogue's scatter) must have landed
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
        if constexpr (kPair) {     // my 64 rows of the tile, into both CTAs (box = 64 rows x 64 columns = 8 KB)
          const int half = static_cast<int>(pair_rank) * (kTileKV / 2);
          tma_load_2d_mcast(dst + half * 128, tm, &kv_full[slot], col, row + half, 0x3);
          tma_load_2d_mcast(dst + kBoxBytes + half * 128, tm, &kv_full[slot], col + 64, row + half, 0x3);
        } else {
          tma_load_2d(dst, tm, &kv_full[slot], col, row);
          tma_load_2d(dst + kBoxBytes, tm, &kv_full[slot], col + 64, row);
        }
        if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
      }
    } else if (warp_idx == 1) {
      // ---------------------------------------------------------------- MMA issuer
      // The WHOLE warp runs this loop in uniform control flow and one elected lane issues: the descriptors then live
      // in uniform registers, which is what UTCHMMA reads (a lane-divergent loop costs an R2UR round trip per operand
      // and made the issue thread, not the tensor pipe, the bottleneck: 80 cycles per MMA measured).
      constexpr uint32_t idesc_qk = make_idesc_bf16(kTileQ, kTileKV, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, kHeadDim, false, true);   // B = V is MN-major
      const bool leader = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_lo = ((smem_u32(smem_q) & 0x3FFFF) >> 4) | kLoKMajor;
      const uint32_t kv_lo = ((smem_u32(smem_kv) & 0x3FFFF) >> 4) | kLoKMajor;
      const uint32_t v_lo = ((smem_u32(smem_kv) & 0x3FFFF) >> 4) | kLoMnMajor;
      int slot = 0;
      uint32_t phase = 0;
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[slot], phase);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (leader) {
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k)
            umma_ss<1>(tm + t * 128, desc64(q_lo + t * (kTileBytes >> 4) + kstep_off(k)),
                       desc64(kv_lo + slot * (kTileBytes >> 4) + kstep_off(k)), idesc_qk, k != 0);
          umma_commit(&s_full[t]);
        }
      }
      if (leader) {
        if constexpr (kPair) umma_commit_mcast(&kv_empty[slot], 0x3); else umma_commit(&kv_empty[slot]);
      }
      __syncwarp();
      if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
      PROF_DECL;
      for (int j = 0; j < n_kv; ++j) {
        const int v_slot = slot;
        PROF(0);
        mbar_wait(&kv_full[v_slot], phase);
        if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
        const bool more = j + 1 < n_kv;
        int k_slot = 0;
        if (more) {
          k_slot = slot;
          mbar_wait(&kv_full[k_slot], phase);
          if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
        }
        tc_fence_after();
        PROF(1);
        const uint32_t v_desc = v_lo + v_slot * (kTileBytes >> 4);
        const uint32_t k_desc = kv_lo + k_slot * (kTileBytes >> 4);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // O_t (+)= P_t V_j : A = P_t (bf16 pairs packed in TMEM columns, 8 columns per k-step of 16),
          //                    B = 16 rows of V (2048 B) x 128 columns (two 64-wide chunks 16 KB apart).
          // P arrives in two halves (keys 0..63, 64..127) so the first four k-steps overlap the rest of the softmax.
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            PROF(2);
            mbar_wait(&p_full[2 * t + half], j & 1);
            tc_fence_after();
            PROF(3 + half);
            if (leader) {
#pragma unroll
              for (int k = half * 4; k < half * 4 + 4; ++k)
                umma_ts(tm + 256 + t * 128, tm + t * 128 + k * 8, desc64(v_desc + k * (2048 >> 4)), idesc_pv, (j | k) != 0);
            }
          }
          if (leader) {
            if (t == 1) {
              if constexpr (kPair) umma_commit_mcast(&kv_empty[v_slot], 0x3); else umma_commit(&kv_empty[v_slot]);
            }
            if (more) {
#pragma unroll
              for (int k = 0; k < kHeadDim / 16; ++k)
                umma_ss<1>(tm + t * 128, desc64(q_lo + t * (kTileBytes >> 4) + kstep_off(k)), desc64(k_desc + kstep_off(k)),
                           idesc_qk, k != 0);
            }
            umma_commit(&s_full[t]);   // S_t(j+1) ready (and P_t V_j done); after the last tile: O_t final
            if (more && t == 1) {
              if constexpr (kPair) umma_commit_mcast(&kv_empty[k_slot], 0x3); else umma_commit(&kv_empty[k_slot]);
            }
          }
          __syncwarp();
        }
      }
      PROF(0);
      PROF_DUMP(8);
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    reg_inc<224>();
    const int t = (warp_idx - 4) >> 2;             // query tile of this warpgroup
    const int quad = warp_idx & 3;                 // TMEM lane quadrant
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + t * 128;
    const uint32_t o_tmem = tmem_base + lane_off + 256 + t * 128;
    // Reference M (log2 domain: P = 2^(s*c - M)) is LAZY and max-free: tile 0 sets it to the exact row maximum; afterwards
    // it only moves when the previous tile's row sum reached 2^8 (checked at the top of the next iteration, when O_t is
    // quiescent), by log2 of that sum — within 2^7 of the true maximum, and any reference both P and l share is exact
    // algebra.  The fast path therefore has no max instruction at all (FMNMX is half rate and the FMA/ALU pipe, not
    // MUFU alone, is what bounds the softmax: tools/ubench), and a tile's P may exceed 2^8 only by the growth inside
    // that one tile, which the fp32 / bf16 exponent absorbs.
    float M = 0.f, l = 0.f, prev_sum = 0.f;
    const bool safe = p.force_safe > 0 ||
                      (p.force_safe == 0 && (p.logit_bound == nullptr || !(__ldg(p.logit_bound) <= kMaxCertifiedLogit)));
    PROF_DECL;
    for (int j = 0; j < n_kv; ++j) {
      PROF(0);
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      PROF(1);
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld32(s_tmem, s0);
      tmem_ld32(s_tmem + 32, s1);
      const int valid = p.kv_len - j * kTileKV;    // >= 128 except on a ragged last tile
      if (safe || j == 0) {
        tmem_ld32(s_tmem + 64, s2);
        tmem_ld32(s_tmem + 96, s3);
        tmem_wait_ld();
        float mx = -INFINITY;
        if (valid >= kTileKV) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 1])));   // ptxas fuses these into 3-input FMNMX
            mx = fmaxf(mx, fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1])));
            mx = fmaxf(mx, fmaxf(__uint_as_float(s2[i]), __uint_as_float(s2[i + 1])));
            mx = fmaxf(mx, fmaxf(__uint_as_float(s3[i]), __uint_as_float(s3[i + 1])));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < valid) mx = fmaxf(mx, __uint_as_float(s0[i]));
            if (32 + i < valid) mx = fmaxf(mx, __uint_as_float(s1[i]));
            if (64 + i < valid) mx = fmaxf(mx, __uint_as_float(s2[i]));
            if (96 + i < valid) mx = fmaxf(mx, __uint_as_float(s3[i]));
          }
        }
        const float m_tile = mx * kScaleLog2;
        if (j == 0) {
          M = m_tile;
        } else {
          // safe: move the reference only when this tile holds a logit more than 2^8 above it (then x <= 8 always)
          const bool need = m_tile - M > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2(M - m_tile) : 1.0f;
            if (need) M = m_tile;
            l *= alpha;
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
      } else {
        const bool grow = !(prev_sum < 256.0f);
        if (__any_sync(0xffffffffu, grow)) {
          // rescale the running sum and this row of O_t (quiescent: P_t V_{j-1} completed before s_full fired)
          const float g = grow ? __log2f(fminf(prev_sum, 3.0e38f)) : 0.f;
          const float alpha = ex2(-g);
          M += g;
          l *= alpha;
          tmem_wait_ld();
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
        tmem_wait_ld();
        tmem_ld32(s_tmem + 64, s2);       // in flight while the first half is exponentiated
        tmem_ld32(s_tmem + 96, s3);
      }
      if (valid < kTileKV) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) s0[i] = 0xff800000u;
          if (32 + i >= valid) s1[i] = 0xff800000u;
        }
      }
      const uint64_t scale2 = pack2(kScaleLog2, kScaleLog2);
      const uint64_t negm2 = pack2(-M, -M);
      uint64_t sum_a = pack2(0.f, 0.f), sum_b = pack2(0.f, 0.f);
      uint32_t pk[32], pk2[32];
      softmax_chunk<kPolyMask>(s0, scale2, negm2, pk, sum_a);
      softmax_chunk<kPolyMask>(s1, scale2, negm2, pk + 16, sum_b);
      tmem_st32(s_tmem, pk);        // P_t columns [0,32): keys 0..63
      PROF(2);
      tmem_wait_ld();
      if (valid < kTileKV) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (64 + i >= valid) s2[i] = 0xff800000u;
          if (96 + i >= valid) s3[i] = 0xff800000u;
        }
      }
      softmax_chunk<kPolyMask>(s2, scale2, negm2, pk2, sum_a);
      PROF(3);
      tmem_wait_st();               // first half landed while chunk 2 was computed: hand it to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * t]);
      PROF(5);
      softmax_chunk<kPolyMask>(s3, scale2, negm2, pk2 + 16, sum_b);
      tmem_st32(s_tmem + 32, pk2);  // P_t columns [32,64): keys 64..127
      PROF(6);
      {
        float a0, a1, b0, b1;
        unpack2(sum_a, a0, a1);
        unpack2(sum_b, b0, b1);
        prev_sum = (a0 + a1) + (b0 + b1);
        l += prev_sum;
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * t + 1]);
      PROF(7);
    }
    if (warp_idx == 4 && lane == 0) PROF_DUMP(0);
    // ---- epilogue: O_t / l -> bf16 -> global (row = one thread, 256 contiguous bytes per head)
    mbar_wait(&s_full[t], n_kv & 1);
    tc_fence_after();
    const int row = q0 + t * kTileQ + quad * 32 + lane;
    __nv_bfloat16* dst = nullptr;
    if (row < p.q_len) {
      const int owner = row / p.rows_per_rank;
      const int b = head / p.heads_per_batch, hl = head - b * p.heads_per_batch;
      dst = static_cast<__nv_bfloat16*>(p.o_peers[owner]) +
            (static_cast<int64_t>(b) * p.batch_rows + row - owner * p.rows_per_rank) * p.ld_o + p.col0 + hl * kHeadDim;
    }
    if (p.ring_o == nullptr) {
      const float inv_l = 1.0f / l;
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
    } else {
      // merge this K/V block into the running state: m' = max(m, M); O' = O 2^(m-m') + O_blk 2^(M-m'); l likewise
      const bool ok = row < p.q_len;
      const int64_t rh = static_cast<int64_t>(ok ? row : 0) * p.num_heads + head;
      float m_old = -INFINITY, l_old = 0.f;
      if (!p.ring_first && ok) {
        const float2 ml = *reinterpret_cast<const float2*>(p.ring_ml + 2 * rh);
        m_old = ml.x;
        l_old = ml.y;
      }
      const float m_new = fmaxf(m_old, M);
      const float a = p.ring_first ? 0.f : ex2(m_old - m_new), b = ex2(M - m_new);
      const float l_new = l_old * a + l * b;
      const float inv_l = 1.0f / l_new;
      float* so = p.ring_o + rh * kHeadDim;
      if (!p.ring_last && ok) *reinterpret_cast<float2*>(p.ring_ml + 2 * rh) = make_float2(m_new, l_new);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t o[32];
        tmem_ld32(o_tmem + c * 32, o);
        tmem_wait_ld();
        if (!ok) continue;
        float v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!p.ring_first) prev = *reinterpret_cast<const float4*>(so + c * 32 + g * 4);
          v[g * 4 + 0] = prev.x * a + __uint_as_float(o[g * 4 + 0]) * b;
          v[g * 4 + 1] = prev.y * a + __uint_as_float(o[g * 4 + 1]) * b;
          v[g * 4 + 2] = prev.z * a + __uint_as_float(o[g * 4 + 2]) * b;
          v[g * 4 + 3] = prev.w * a + __uint_as_float(o[g * 4 + 3]) * b;
        }
        if (p.ring_last) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = pack_bf16x2(v[g * 8 + 2 * i] * inv_l, v[g * 8 + 2 * i + 1] * inv_l);
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(so + c * 32 + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
      }
    }
    if (p.sync.signal_epoch != 0) {   // uniform: every output row of this CTA is stored -> count it; the last CTA tells the peers
      __threadfence_system();
      named_bar_sync(1, 256);
      if (warp_idx == 4 && lane == 0) cp_signal_when_grid_done(p.sync, gridDim.x * gridDim.y);
    }
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (kPair) cluster_sync(); else __syncthreads();   // no CTA leaves while its peer can still write into it
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}


}  // namespace
}  // namespace drb

namespace drb {
namespace {
struct AttnLaunchExtra {
  float* ring_o = nullptr;
  float* ring_ml = nullptr;
  int ring_first = 0, ring_last = 0;
  int heads_per_batch = 0, batch_rows = 0;   // 0: unbatched
  const float* logit_bound = nullptr;        // the caller's certificate (device float); NULL = none (safe softmax)
  const drb_cp_sync* sync = nullptr;
};

template <uint32_t kMask>
int attention_configure() {
  DRB_CUDA(cudaFuncSetAttribute(attention_kernel<kMask, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  DRB_CUDA(cudaFuncSetAttribute(attention_kernel<kMask, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  return 0;
}

template <uint32_t kMask>
int attention_dispatch(const cudaLaunchConfig_t& cfg, bool pair, const CUtensorMap& tq, const CUtensorMap& tk,
                       const CUtensorMap& tv, const AttnParams& p) {
  static DeviceOnce configured;   // per mask and per device: the attribute belongs to the device's context
  const int rc = device_once(configured, [] { return attention_configure<kMask>(); });
  if (rc) return rc;
  if (pair) DRB_CUDA(cudaLaunchKernelEx(&cfg, attention_kernel<kMask, true>, tq, tk, tv, p));
  else DRB_CUDA(cudaLaunchKernelEx(&cfg, attention_kernel<kMask, false>, tq, tk, tv, p));
  return 0;
}
}  // namespace
}  // namespace drb

static int attention_launch(const void* q, const void* k, const void* v, int64_t ld_qkv, void* const* o_peers, int world,
                            int64_t ld_o, int q_len, int kv_len, int num_heads, int rows_per_rank, int col0, void* stream,
                            const drb::AttnLaunchExtra& ex = drb::AttnLaunchExtra()) {
  using namespace drb;
  DRB_REQUIRE(q && k && v && o_peers, "null pointer");
  DRB_REQUIRE(q_len > 0 && kv_len > 0 && num_heads > 0, "q_len, kv_len, num_heads must be positive");
  DRB_REQUIRE(ld_qkv % 8 == 0 && ld_o % 8 == 0 && col0 % 8 == 0 && col0 >= 0, "row pitches / column offset must be multiples of 8 elements");
  const int hpb = ex.heads_per_batch > 0 ? ex.heads_per_batch : num_heads;
  DRB_REQUIRE(num_heads % hpb == 0 && ex.batch_rows >= 0, "heads_per_batch must divide the head count");
  DRB_REQUIRE(ld_qkv >= static_cast<int64_t>(num_heads) * kHeadDim && ld_o >= col0 + static_cast<int64_t>(hpb) * kHeadDim,
              "row pitch smaller than the heads it holds");
  DRB_REQUIRE(world >= 1 && world <= DRB_CP_MAX_RANKS && rows_per_rank > 0 &&
                  static_cast<int64_t>(rows_per_rank) * world >= q_len, "rows_per_rank * world must cover q_len");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(ex.logit_bound) & 3) == 0, "logit bound pointer misaligned");
  AttnParams p{};
  {
    const int rc_sync = fill_cp_sync(&p.sync, ex.sync);
    if (rc_sync) return rc_sync;
  }
  for (int i = 0; i < world; ++i) {
    DRB_REQUIRE(o_peers[i] != nullptr && (reinterpret_cast<uintptr_t>(o_peers[i]) & 15) == 0, "o not 16-byte aligned");
    p.o_peers[i] = o_peers[i];
  }
  p.ld_o = ld_o;
  p.q_len = q_len;
  p.kv_len = kv_len;
  p.rows_per_rank = rows_per_rank;
  p.col0 = col0;
  p.heads_per_batch = hpb;
  p.batch_rows = ex.batch_rows;
  p.ring_o = ex.ring_o;
  p.ring_ml = ex.ring_ml;
  p.ring_first = ex.ring_first;
  p.ring_last = ex.ring_last;
  p.num_heads = num_heads;
  // DRB_ATTN_PAIR (0 = no K/V multicast) and DRB_ATTN_SAFE (1 = always the per-tile-max softmax, 0 = never: tuning only)
  // are A/B switches; the fraction of exponentials on the FMA pipe is fixed at 5/16 (the other fractions
  // are compiled only with -DDRB_ATTN_TUNE and selected by DRB_ATTN_POLY).
  static const int pair_ok = [] { const char* e = getenv("DRB_ATTN_PAIR"); return e ? atoi(e) : 1; }();
  static const int force_safe = [] { const char* e = getenv("DRB_ATTN_SAFE"); return e ? atoi(e) : -1; }();
  p.logit_bound = ex.logit_bound;
  p.force_safe = force_safe < 0 ? 0 : (force_safe != 0 ? 1 : -1);
  dim3 grid((q_len + kTileQ * kQTilesPerCta - 1) / (kTileQ * kQTilesPerCta), num_heads);
  const bool pair = pair_ok && (grid.x % 2 == 0);       // clusters of two adjacent query blocks of one head
  CUtensorMap tq, tk, tv;
  const uint64_t cols = static_cast<uint64_t>(num_heads) * kHeadDim;
  int rc = make_tmap_2d_bf16(&tq, q, q_len, cols, ld_qkv, kTileQ, 64);
  if (rc) return rc;
  // pair: K/V maps with 64-row boxes — each CTA of the pair loads (and multicasts) half of every tile
  rc = make_tmap_2d_bf16(&tk, k, kv_len, cols, ld_qkv, pair ? kTileKV / 2 : kTileKV, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tv, v, kv_len, cols, ld_qkv, pair ? kTileKV / 2 : kTileKV, 64);
  if (rc) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = kAttnSmem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#ifdef DRB_ATTN_TUNE
  static const int variant = [] { const char* e = getenv("DRB_ATTN_POLY"); return e ? atoi(e) : 2; }();
  switch (variant) {   // 0 -> 0/16, 1 -> 4/16, 2 -> 5/16, 3 -> 8/16, 4 -> 3/16, 5 -> 2/16, 6 -> 6/16, 7 -> 7/16
    case 0: return attention_dispatch<0x0000u>(cfg, pair, tq, tk, tv, p);
    case 1: return attention_dispatch<0x1111u>(cfg, pair, tq, tk, tv, p);
    case 3: return attention_dispatch<0x5555u>(cfg, pair, tq, tk, tv, p);
    case 4: return attention_dispatch<0x0421u>(cfg, pair, tq, tk, tv, p);
    case 5: return attention_dispatch<0x0101u>(cfg, pair, tq, tk, tv, p);
    case 6: return attention_dispatch<0x2929u>(cfg, pair, tq, tk, tv, p);
    case 7: return attention_dispatch<0x52A5u>(cfg, pair, tq, tk, tv, p);
    default: break;
  }
#endif
  return attention_dispatch<0x4924u>(cfg, pair, tq, tk, tv, p);
}

extern "C" int drb_attention_bf16(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                                  int q_len, int kv_len, int num_heads, void* stream) {
  return drb_attention_bf16_bounded(q, k, v, ld_qkv, o, ld_o, q_len, kv_len, num_heads, nullptr, stream);
}

extern "C" int drb_attention_bf16_bounded(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                                          int q_len, int kv_len, int num_heads, const float* max_abs_logit, void* stream) {
  void* peers[1] = {o};
  drb::AttnLaunchExtra ex;
  ex.logit_bound = max_abs_logit;
  return attention_launch(q, k, v, ld_qkv, peers, 1, ld_o, q_len, kv_len, num_heads, 0x7fffffff, 0, stream, ex);
}

extern "C" int drb_attention_bf16_cp(const void* q, const void* k, const void* v, int64_t ld_qkv, void* const* o_peers, int world,
                                     int64_t ld_o, int q_len, int kv_len, int num_heads, int rows_per_rank, int col0,
                                     void* stream) {
  return drb_attention_bf16_cp_batched(q, k, v, ld_qkv, o_peers, world, ld_o, q_len, kv_len, num_heads, rows_per_rank, col0, num_heads,
                                       0, nullptr, nullptr, stream);
}

extern "C" int drb_attention_bf16_cp_batched(const void* q, const void* k, const void* v, int64_t ld_qkv, void* const* o_peers,
                                             int world, int64_t ld_o, int q_len, int kv_len, int num_heads, int rows_per_rank,
                                             int col0, int heads_per_batch, int batch_rows, const float* max_abs_logit,
                                             const drb_cp_sync* sync, void* stream) {
  drb::AttnLaunchExtra ex;
  ex.heads_per_batch = heads_per_batch;
  ex.batch_rows = batch_rows;
  ex.logit_bound = max_abs_logit;
  ex.sync = sync;
  return attention_launch(q, k, v, ld_qkv, o_peers, world, ld_o, q_len, kv_len, num_heads, rows_per_rank, col0, stream, ex);
}

extern "C" int drb_attention_bf16_ring(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                                       float* state_o, float* state_ml, int q_len, int kv_len, int num_heads, int first, int last,
                                       void* stream) {
  return drb_attention_bf16_ring_bounded(q, k, v, ld_qkv, o, ld_o, state_o, state_ml, q_len, kv_len, num_heads, first, last, nullptr,
                                         stream);
}

extern "C" int drb_attention_bf16_ring_bounded(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                                               float* state_o, float* state_ml, int q_len, int kv_len, int num_heads, int first,
                                               int last, const float* max_abs_logit, void* stream) {
  using namespace drb;
  DRB_REQUIRE(state_o && state_ml, "ring attention needs the running-state buffers");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(state_o) & 15) == 0 && (reinterpret_cast<uintptr_t>(state_ml) & 7) == 0, "state buffers misaligned");
  DRB_REQUIRE(!last || o != nullptr, "the last block writes the output");
  void* peers[1] = {o ? o : static_cast<void*>(state_o)};
  AttnLaunchExtra ex;
  ex.ring_o = state_o;
  ex.ring_ml = state_ml;
  ex.ring_first = first ? 1 : 0;
  ex.ring_last = last ? 1 : 0;
  ex.logit_bound = max_abs_logit;
  return attention_launch(q, k, v, ld_qkv, peers, 1, o ? ld_o : static_cast<int64_t>(num_heads) * kHeadDim, q_len, kv_len, num_heads,
                          0x7fffffff, 0, stream, ex);
}

#ifdef DRB_ATTN_PROFILE
using namespace drb;
extern "C" int drb_debug_attn_profile(unsigned long long* out16) {
  return drb::check_cuda(cudaMemcpyFromSymbol(out16, g_attn_prof, sizeof(unsigned long long) * 16), "drb_debug_attn_profile");
}
#endif
