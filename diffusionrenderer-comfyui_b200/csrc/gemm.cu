// drb_gemm_bf16 — out[M,N] = epilogue(A[M,K] @ W[N,K]^T) on tcgen05 tensor cores.
//
// Replaces every nn.Linear on the token stream of the reference DiT (CleanGeneralDIT.py:273-276, :304, :454-460,
// :417, :590).  bf16 operands staged by TMA (128-byte swizzle), fp32 accumulators in TMEM, bf16 output.
//
// Structure (persistent, warp-specialised, one CTA — or one CTA pair — per SM):
//   warp 0      TMA producer            (ring of kStages {A,B} k-blocks of 64)
//   warp 1      tcgen05.mma issuer      (one thread; leader CTA only in pair mode)
//   warp 2      TMEM allocator          (512 columns = 2 accumulator stages x 256 fp32 columns)
//   warps 4..7  epilogue                (tcgen05.ld -> registers -> epilogue math -> swizzled smem staging -> coalesced row stores)
// Epilogues: store, exact-erf GELU, gated residual, and the QKV projection's (per-head RMSNorm + RoPE in registers, rows stored
// locally or — under context parallelism — straight into the GPU that owns the head: drb_gemm_qkv_norm_rope).
// kCtaGroup = 1: tile 128 x 256 per CTA.  kCtaGroup = 2: tile 256 x 256 per CTA pair (cta_group::2 MMA, each CTA
// stages its own 128 rows of A and its own 128-row half of W, halving the shared-memory/L2 operand traffic).
// The MMA of tile i+1 overlaps the epilogue of tile i through the two TMEM accumulator stages.
//
// Roofline: tensor pipe.  Algorithmic work = 2*M*N*K flop per launch.
#include <math.h>
#include <stdlib.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "cp_sync.cuh"
#include "ptx.cuh"

namespace drb {
namespace {

constexpr int kBlockM = 128;   // rows of A per CTA
constexpr int kBlockN = 256;   // columns of the output tile (rows of W)
constexpr int kBlockK = 64;    // one 128-byte swizzle atom of bf16
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 256;
constexpr int kMaxBandTilesN = 16;   // rasterisation: bands of <= 16 n-tiles whose W rows (band * 256 * K * 2 B) stay in L2

template <int kCtaGroup>
struct GemmCfg {
  static constexpr int kBRows = kBlockN / kCtaGroup;                      // W rows staged per CTA
  static constexpr int kABytes = kBlockM * kBlockK * 2;                   // 16 KB
  static constexpr int kBBytes = kBRows * kBlockK * 2;                    // 32 KB / 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;                   // 48 KB / 32 KB
  static constexpr int kStages = kCtaGroup == 1 ? 4 : 6;                  // 192 KB of operand ring
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  // the QKV epilogue stages one head per warp (32 rows x 256 B, 16-byte chunks XOR-swizzled by row) so that whole rows
  // go out coalesced; only that instantiation allocates it
  static constexpr int kEpiStageBytes = 4 * 32 * 256;
};

struct GemmParams {
  int M, N, K;
  int band;      // n-tiles per rasterisation band
  int stage_out; // generic epilogues: stage each 32-column chunk through shared memory and store 64-byte row segments
  __nv_bfloat16* out;
  int64_t ldo;
  const __nv_bfloat16* resid;
  int64_t ldr;
  const __nv_bfloat16* gate;
  // DRB_EPI_QKV_NORM_ROPE only: N = 3*sect (q | k | v), per-head RMSNorm weights, RoPE tables [M,128]; with world > 0 the
  // rows go to the GPU that owns the head (context parallelism: the Ulysses exchange is this epilogue's store)
  const __nv_bfloat16* wq;
  const __nv_bfloat16* wk;
  const __nv_bfloat16* cos_tab;
  const __nv_bfloat16* sin_tab;
  int sect;
  int world, heads_per_rank, row0;
  int64_t peer_ld;
  void* peers[DRB_CP_MAX_RANKS];
  // batched sequences along M (the five G-buffer passes / cond + uncond of one video): row r belongs to sequence
  // r / rows_per_batch; under context parallelism sequence b's heads occupy columns (b*heads_per_rank + h)*128 of each
  // section of the peer row, sections being `batch` times wider.  batch == 1: rows_per_batch = INT_MAX / 2.
  int rows_per_batch, batch;
  // context parallelism: wait for the peers' stores into A before the first load / signal the peers once every CTA has
  // stored its rows (csrc/cp_sync.cuh); both off unless a drb_cp_sync descriptor was passed
  CpSync sync;
};

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int band_n, int& m, int& n) {
  const int per_band = tiles_m * band_n;
  const int band = t / per_band;
  const int r = t - band * per_band;
  const int n0 = band * band_n;
  const int w = min(band_n, tiles_n - n0);
  m = r / w;
  n = n0 + (r - m * w);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int kCtaGroup, int kEpi>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<kCtaGroup>;
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_stage = smem + Cfg::kStages * Cfg::kStageBytes + 256;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCtaGroup == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;

  const int tile_m_rows = kBlockM * kCtaGroup;
  const int tiles_m = (p.M + tile_m_rows - 1) / tile_m_rows;
  const int tiles_n = (p.N + kBlockN - 1) / kBlockN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int cluster_id = blockIdx.x / kCtaGroup;
  const int num_clusters = gridDim.x / kCtaGroup;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4 * kCtaGroup);   // one arrival per epilogue warp of every CTA in the group
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc<kCtaGroup>(tmem_slot, 512);
    tmem_relinquish<kCtaGroup>();
  }
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      cp_wait(p.sync);   // A rows stored by peer GPUs (the attention epilogue's scatter) must have landed
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int tm, tn;
        tile_coords(t, tiles_m, tiles_n, p.band, tm, tn);
        const int row_a = tm * tile_m_rows + static_cast<int>(cta_rank) * kBlockM;
        const int row_b = tn * kBlockN + static_cast<int>(cta_rank) * Cfg::kBRows;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if constexpr (kCtaGroup == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, row_a);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBlockK, row_b);
          } else {
            // both CTAs' bytes are accounted on the leader's barrier
            if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * 2);
            tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * kBlockK, row_a);
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kb * kBlockK, row_b);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The whole warp runs the loop in uniform control flow and one elected lane issues, so that the descriptors
    // live in uniform registers (what UTCHMMA reads) instead of paying an R2UR round trip per operand.
    if (is_leader) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM * kCtaGroup, kBlockN);
      const bool issuer = elect_one();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_lo = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);                   // K-major, LBO field = 1
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);                        // SBO 1024 B, v1, SWIZZLE_128B
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);   // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tm + as * kBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = smem_lo + stage * (Cfg::kStageBytes >> 4);
          const uint32_t b_lo = a_lo + (Cfg::kABytes >> 4);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advance 32 bytes (16 bf16) inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
              umma_ss<kCtaGroup>(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k),
                                 (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k), idesc, (kb | k) != 0);
            }
            if constexpr (kCtaGroup == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_pair(&empty_bar[stage], 0x3);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (issuer) {
          if constexpr (kCtaGroup == 1) umma_commit(&tmem_full_bar[as]);
          else umma_commit_pair(&tmem_full_bar[as], 0x3);
        }
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    const int q = warp_idx - 4;   // == warp_idx % 4: the TMEM lane quadrant this warp may read
    int iter = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
      int tm, tn;
      tile_coords(t, tiles_m, tiles_n, p.band, tm, tn);
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const int row = tm * tile_m_rows + static_cast<int>(cta_rank) * kBlockM + q * 32 + lane;
      const int col0 = tn * kBlockN;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kBlockN;
      const bool row_ok = row < p.M;
      __nv_bfloat16* out_row = p.out + static_cast<int64_t>(row) * p.ldo;
      const __nv_bfloat16* res_row = kEpi == DRB_EPI_GATED_RESIDUAL ? p.resid + static_cast<int64_t>(row) * p.ldr : nullptr;
      if constexpr (kEpi == DRB_EPI_QKV_NORM_ROPE) {
        // The 256-column tile is two whole heads of one section (q, k or v).  One thread owns one token row, so the
        // per-head RMSNorm and the rotate-half partner (d <-> d +- 64) are thread-local: bf16(acc) -> norm -> RoPE with
        // the roundings of CleanGeneralDIT.py:23-33, :67-80 (same arithmetic as qk_norm_rope_kernel).
        const int sect = col0 / p.sect;
        const int head0 = (col0 - sect * p.sect) >> 7;
        const __nv_bfloat16* wn = sect == 1 ? p.wk : p.wq;
        const __nv_bfloat16* crow = p.cos_tab + static_cast<int64_t>(row_ok ? row : 0) * 128;
        const __nv_bfloat16* srow = p.sin_tab + static_cast<int64_t>(row_ok ? row : 0) * 128;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          if (col0 + hh * 128 >= p.N) break;   // warp-uniform
          uint32_t r[4][32];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld32(taddr + hh * 128 + c * 32, r[c]);
          tmem_wait_ld();
          float inv = 1.0f;
          if (sect < 2) {
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = bf16_round(__uint_as_float(r[c][i]));
                r[c][i] = __float_as_uint(x);
                ss += x * x;
              }
            inv = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wn + c * 32 + g * 8));
                const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  r[c][g * 8 + 2 * j] = __float_as_uint(bf16_round(__uint_as_float(r[c][g * 8 + 2 * j]) * inv * bf16_lo(ww[j])));
                  r[c][g * 8 + 2 * j + 1] = __float_as_uint(bf16_round(__uint_as_float(r[c][g * 8 + 2 * j + 1]) * inv * bf16_hi(ww[j])));
                }
              }
          }
          // RoPE (q, k) or plain rounding (v) into this warp's staging rows, then whole 256-byte head rows go out
          // coalesced (two rows per store instruction) — to the local buffer, or over NVLink to the head's owner
          uint8_t* my_stage = epi_stage + q * (32 * 256);
          uint8_t* my_row = my_stage + lane * 256;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t o[4];
              if (sect < 2) {
                const uint4 cv = *reinterpret_cast<const uint4*>(crow + c * 32 + g * 8);
                const uint4 sv = *reinterpret_cast<const uint4*>(srow + c * 32 + g * 8);
                const uint32_t cc[4] = {cv.x, cv.y, cv.z, cv.w}, sn[4] = {sv.x, sv.y, sv.z, sv.w};
                const float sign = c < 2 ? -1.0f : 1.0f;           // rotate_half = cat(-x[64:], x[:64])
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int i0 = g * 8 + 2 * j;
                  const float a0 = bf16_round(__uint_as_float(r[c][i0]) * bf16_lo(cc[j]));
                  const float a1 = bf16_round(__uint_as_float(r[c][i0 + 1]) * bf16_hi(cc[j]));
                  const float b0 = bf16_round(sign * __uint_as_float(r[c ^ 2][i0]) * bf16_lo(sn[j]));
                  const float b1 = bf16_round(sign * __uint_as_float(r[c ^ 2][i0 + 1]) * bf16_hi(sn[j]));
                  o[j] = pack_bf16x2(a0 + b0, a1 + b1);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  o[j] = pack_bf16x2(__uint_as_float(r[c][g * 8 + 2 * j]), __uint_as_float(r[c][g * 8 + 2 * j + 1]));
              }
              *reinterpret_cast<uint4*>(my_row + (((c * 4 + g) ^ (lane & 15)) * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          __syncwarp();
          const int head = head0 + hh;
          const int row_base = row - lane;                          // first row of this warp's 32
          if (p.world > 0) {
            // destination row = row0 + (token index inside its sequence); the 32 rows of a warp straddle at most one
            // sequence boundary (rows_per_batch >= 32), so the sequence index is b0 or b0 + 1
            const int owner = head / p.heads_per_rank;
            const int b0 = row_base / p.rows_per_batch;
            const int bound = (b0 + 1) * p.rows_per_batch;
            __nv_bfloat16* dst0 = static_cast<__nv_bfloat16*>(p.peers[owner]) +
                                  (static_cast<int64_t>(sect) * p.batch + b0) * p.heads_per_rank * 128 +
                                  (head - owner * p.heads_per_rank) * 128 + (lane & 15) * 8;
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
              const int rr = it * 2 + (lane >> 4);
              const int grow = row_base + rr;
              const int nb = grow >= bound ? 1 : 0;
              if (grow < p.M)
                *reinterpret_cast<uint4*>(dst0 + static_cast<int64_t>(p.row0 + grow - (b0 + nb) * p.rows_per_batch) * p.peer_ld +
                                          nb * p.heads_per_rank * 128) =
                    *reinterpret_cast<const uint4*>(my_stage + rr * 256 + (((lane & 15) ^ (rr & 15)) * 16));
            }
          } else {
            __nv_bfloat16* dst0 = p.out + static_cast<int64_t>(row_base) * p.ldo + col0 + hh * 128;
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
              const int rr = it * 2 + (lane >> 4);
              if (row_base + rr < p.M)
                *reinterpret_cast<uint4*>(dst0 + rr * p.ldo + (lane & 15) * 8) =
                    *reinterpret_cast<const uint4*>(my_stage + rr * 256 + (((lane & 15) ^ (rr & 15)) * 16));
            }
          }
          __syncwarp();
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < kBlockN / 32; ++c) {
        const int col = col0 + c * 32;
        if (col >= p.N) break;   // warp-uniform
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // 4 groups of 8 columns = one 16-byte store each
          const int cg = col + g * 8;
          if (!row_ok || cg >= p.N) continue;
          uint32_t o[4];
          if constexpr (kEpi == DRB_EPI_STORE) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = pack_bf16x2(__uint_as_float(r[g * 8 + 2 * j]), __uint_as_float(r[g * 8 + 2 * j + 1]));
          } else if constexpr (kEpi == DRB_EPI_GELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a0 = bf16_round(__uint_as_float(r[g * 8 + 2 * j]));
              const float a1 = bf16_round(__uint_as_float(r[g * 8 + 2 * j + 1]));
              o[j] = pack_bf16x2(gelu_erf(a0), gelu_erf(a1));
            }
          } else {
            const uint4 rv = *reinterpret_cast<const uint4*>(res_row + cg);
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(p.gate + cg));
            const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
            const uint32_t gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a0 = bf16_round(__uint_as_float(r[g * 8 + 2 * j]));
              const float a1 = bf16_round(__uint_as_float(r[g * 8 + 2 * j + 1]));
              const float t0 = bf16_round(bf16_lo(gg[j]) * a0);
              const float t1 = bf16_round(bf16_hi(gg[j]) * a1);
              o[j] = pack_bf16x2(bf16_lo(rr[j]) + t0, bf16_hi(rr[j]) + t1);
            }
          }
          if (p.stage_out)
            *reinterpret_cast<uint4*>(epi_stage + q * 2048 + lane * 64 + ((g ^ ((lane >> 1) & 3)) * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
          else
            *reinterpret_cast<uint4*>(out_row + cg) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (p.stage_out) {
          // rows of this warp, 64 bytes (32 columns) each: 8 rows per store instruction, 4 lanes per row
          __syncwarp();
          const int row_base = row - lane;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + (lane >> 2), ch = lane & 3;
            if (row_base + rr < p.M && col + ch * 8 < p.N)
              *reinterpret_cast<uint4*>(p.out + static_cast<int64_t>(row_base + rr) * p.ldo + col + ch * 8) =
                  *reinterpret_cast<const uint4*>(epi_stage + q * 2048 + rr * 64 + ((ch ^ ((rr >> 1) & 3)) * 16));
          }
          __syncwarp();
        }
      }
      }
      // release the accumulator stage to the MMA issuer (the leader's barrier)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtaGroup == 1) mbar_arrive(&tmem_empty_bar[as]);
        else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      }
    }
    if (p.sync.signal_epoch != 0) {   // uniform: every row of this CTA is stored -> count it; the last CTA tells the peers
      __threadfence_system();
      named_bar_sync(1, 128);
      if (warp_idx == 4 && lane == 0) cp_signal_when_grid_done(p.sync, gridDim.x);
    }
  }

  // ------------------------------------------------------------------ teardown
  __syncwarp();   // reconverge the single-lane roles before the aligned barriers below
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<kCtaGroup>(tmem_base, 512);
  }
}

template <int kCtaGroup, int kEpi>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<kCtaGroup>;
  auto kernel = gemm_bf16_kernel<kCtaGroup, kEpi>;
  constexpr int kSmem = Cfg::kSmemBytes + (kEpi == DRB_EPI_QKV_NORM_ROPE ? Cfg::kEpiStageBytes : 4 * 2048);
  static_assert(kSmem <= 232448, "over the 227 KB shared-memory limit of sm_100");
  static DeviceOnce configured;   // per template instance and per device (the attribute belongs to the device's context)
  {
    const int rc = device_once(configured, [&] {
      return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem), "gemm_bf16_kernel smem");
    });
    if (rc) return rc;
  }
  const int tile_m_rows = kBlockM * kCtaGroup;
  const int tiles = ((p.M + tile_m_rows - 1) / tile_m_rows) * ((p.N + kBlockN - 1) / kBlockN);
  int sms = num_sms();
  int clusters = sms / kCtaGroup;
  if (clusters > tiles) clusters = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCtaGroup);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtaGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DRB_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta, tb, p));
  return 0;
}

// n-tiles per band: the W rows of a band (256 * K * 2 B per n-tile) should sit in L2 (126 MB) next to the streaming A rows
int pick_band(int K) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("DRB_GEMM_BAND");      // tuning runs only
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) return forced;
  const int64_t per_tile = 256LL * K * 2;
  int band = static_cast<int>((64LL << 20) / per_tile);   // K = 16384 -> 8 (measured best: 1511 vs 1487 TF/s at 16), K <= 8192 -> 16
  if (band < 1) band = 1;
  if (band > kMaxBandTilesN) band = kMaxBandTilesN;
  return band;
}

template <int kCtaGroup>
int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case DRB_EPI_STORE: return launch_gemm<kCtaGroup, DRB_EPI_STORE>(ta, tb, p, s);
    case DRB_EPI_GELU: return launch_gemm<kCtaGroup, DRB_EPI_GELU>(ta, tb, p, s);
    case DRB_EPI_GATED_RESIDUAL: return launch_gemm<kCtaGroup, DRB_EPI_GATED_RESIDUAL>(ta, tb, p, s);
    // DRB_EPI_QKV_NORM_ROPE needs the norm weights / RoPE tables of drb_gemm_qkv_norm_rope: not reachable from here
    default: return fail("drb_gemm_bf16", "epilogue must be DRB_EPI_STORE, DRB_EPI_GELU or DRB_EPI_GATED_RESIDUAL");
  }
}

}  // namespace
}  // namespace drb

extern "C" int drb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M,
                             int N, int K, int epilogue, const void* resid, int64_t ldr, const void* gate,
                             int cta_group, void* stream) {
  return drb_gemm_bf16_sync(A, lda, W, ldw, out, ldo, M, N, K, epilogue, resid, ldr, gate, cta_group, nullptr, stream);
}

extern "C" int drb_gemm_bf16_sync(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M,
                                  int N, int K, int epilogue, const void* resid, int64_t ldr, const void* gate,
                                  int cta_group, const drb_cp_sync* sync, void* stream) {
  using namespace drb;
  DRB_REQUIRE(A && W && out, "null pointer");
  DRB_REQUIRE(M > 0 && N > 0 && K > 0, "M, N, K must be positive");
  DRB_REQUIRE(N % 8 == 0 && K % 8 == 0, "N and K must be multiples of 8");
  DRB_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0, "row pitches must be multiples of 8 elements");
  DRB_REQUIRE(lda >= K && ldw >= K && ldo >= N, "row pitch smaller than the row");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out not 16-byte aligned");
  if (epilogue == DRB_EPI_GATED_RESIDUAL) {
    DRB_REQUIRE(resid && gate, "gated-residual epilogue needs resid and gate");
    DRB_REQUIRE(ldr % 8 == 0 && ldr >= N, "bad residual pitch");
    DRB_REQUIRE((reinterpret_cast<uintptr_t>(resid) & 15) == 0 && (reinterpret_cast<uintptr_t>(gate) & 15) == 0,
                "resid/gate not 16-byte aligned");
  }
  DRB_REQUIRE(cta_group >= 0 && cta_group <= 2, "cta_group must be 0, 1 or 2");
  if (cta_group == 0) cta_group = (M > 128) ? 2 : 1;
  CUtensorMap ta, tb;
  int rc = make_tmap_2d_bf16(&ta, A, M, K, lda, kBlockM, kBlockK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, W, N, K, ldw, kBlockN / cta_group, kBlockK);
  if (rc) return rc;
  GemmParams p{};
  rc = fill_cp_sync(&p.sync, sync);
  if (rc) return rc;
  p.M = M; p.N = N; p.K = K;
  p.band = pick_band(K);
  {
    static int stage = -1;
    if (stage < 0) {
      const char* e = getenv("DRB_GEMM_STAGE");   // tuning runs only
      stage = e ? atoi(e) : 1;
    }
    p.stage_out = stage;
  }
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.resid = static_cast<const __nv_bfloat16*>(resid);
  p.ldr = ldr;
  p.gate = static_cast<const __nv_bfloat16*>(gate);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return cta_group == 1 ? dispatch_epi<1>(epilogue, ta, tb, p, s) : dispatch_epi<2>(epilogue, ta, tb, p, s);
}

extern "C" int drb_gemm_qkv_norm_rope(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M, int D,
                                      int K, const void* wq, const void* wk, const void* cos_tab, const void* sin_tab,
                                      void* const* peer_ptrs, int world, int64_t peer_ld, int row0, void* stream) {
  return drb_gemm_qkv_norm_rope_batched(A, lda, W, ldw, out, ldo, M, D, K, wq, wk, cos_tab, sin_tab, peer_ptrs, world, peer_ld, row0,
                                        1, M, nullptr, stream);
}

extern "C" int drb_gemm_qkv_norm_rope_batched(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M,
                                              int D, int K, const void* wq, const void* wk, const void* cos_tab,
                                              const void* sin_tab, void* const* peer_ptrs, int world, int64_t peer_ld, int row0,
                                              int batch, int rows_per_batch, const drb_cp_sync* sync, void* stream) {
  using namespace drb;
  DRB_REQUIRE(batch >= 1 && rows_per_batch >= 1 && static_cast<int64_t>(batch) * rows_per_batch == M,
              "batch * rows_per_batch must equal M");
  DRB_REQUIRE(batch == 1 || rows_per_batch >= 32, "batched sequences need at least 32 rows each");
  DRB_REQUIRE(A && W && wq && wk && cos_tab && sin_tab, "null pointer");
  DRB_REQUIRE(M > 0 && D > 0 && K > 0 && D % 256 == 0 && K % 8 == 0, "D must be a multiple of 256 (two heads per tile), K of 8");
  DRB_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "row pitches must be multiples of 8 elements and cover K");
  DRB_REQUIRE(((reinterpret_cast<uintptr_t>(wq) | reinterpret_cast<uintptr_t>(wk) | reinterpret_cast<uintptr_t>(cos_tab) |
                reinterpret_cast<uintptr_t>(sin_tab)) & 15) == 0, "norm weights / RoPE tables must be 16-byte aligned");
  GemmParams p{};
  {
    const int rc_sync = fill_cp_sync(&p.sync, sync);
    if (rc_sync) return rc_sync;
  }
  p.M = M; p.N = 3 * D; p.K = K;
  p.band = pick_band(K);
  p.wq = static_cast<const __nv_bfloat16*>(wq);
  p.wk = static_cast<const __nv_bfloat16*>(wk);
  p.cos_tab = static_cast<const __nv_bfloat16*>(cos_tab);
  p.sin_tab = static_cast<const __nv_bfloat16*>(sin_tab);
  p.sect = D;
  p.batch = batch;
  p.rows_per_batch = batch == 1 ? 0x3fffffff : rows_per_batch;
  if (world > 0) {
    DRB_REQUIRE(peer_ptrs != nullptr && world <= DRB_CP_MAX_RANKS && (D / 128) % world == 0 && row0 >= 0, "bad context-parallel arguments");
    DRB_REQUIRE(peer_ld % 8 == 0 && peer_ld >= 3LL * batch * D / world, "peer row pitch too small");
    for (int i = 0; i < world; ++i) {
      DRB_REQUIRE(peer_ptrs[i] != nullptr && (reinterpret_cast<uintptr_t>(peer_ptrs[i]) & 15) == 0, "bad peer pointer");
      p.peers[i] = peer_ptrs[i];
    }
    p.world = world;
    p.heads_per_rank = D / 128 / world;
    p.row0 = row0;
    p.peer_ld = peer_ld;
  } else {
    DRB_REQUIRE(out != nullptr && ldo % 8 == 0 && ldo >= 3LL * D && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "bad output buffer");
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
  }
  const int cta_group = (M > 128) ? 2 : 1;
  CUtensorMap ta, tb;
  int rc = make_tmap_2d_bf16(&ta, A, M, K, lda, kBlockM, kBlockK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, W, 3 * D, K, ldw, kBlockN / cta_group, kBlockK);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return cta_group == 1 ? launch_gemm<1, DRB_EPI_QKV_NORM_ROPE>(ta, tb, p, s) : launch_gemm<2, DRB_EPI_QKV_NORM_ROPE>(ta, tb, p, s);
}
