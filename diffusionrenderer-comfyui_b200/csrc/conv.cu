// drb_conv3d_cl — the tokenizer's causal 3-D convolutions as an implicit GEMM on tcgen05 tensor cores.
//
// Replaces CosmosCausalConv3d of diffusers.AutoencoderKLCosmos (the arithmetic behind CleanVAE.py:50-51,59-60; SURVEY.md
// Appendix B): factorised (1,3,3) spatial and (3,1,1) causal temporal convolutions, 1x1x1 projections, the strided
// convolutions of the down-sampler and the nearest-neighbour-upsample + (1,3,3) convolution of the up-sampler.
// Activations are channels-last per frame, [T][H][W][C] bf16; weights [Cout][kt][kh][kw][Cin] bf16.
//
// out[t, oh, ow, :] = bias + residual + sum over taps (dt,dy,dx) of W[:, tap, :] . x[tsrc(t,dt), s*h + dy - ph, s*w + dx - pw, :]
//
// Implicit GEMM: the M tile is 128 output positions = an 8 x 16 patch (h, w) of one frame; for every tap and every
// 64-channel chunk the TMA producer loads the *shifted* patch with one 4-D box (c, w, h, t) — out-of-bounds rows and
// columns arrive as zeros, which IS the spatial zero padding (the halo is never materialised: each tap is its own box
// served from L2), and the causal replicate-first-frame padding is a clamp of the frame coordinate.  Spatial stride 2
// uses four tensor maps, one per (row, column) parity of the source, so every tap is still a dense box.  The smem tile
// has exactly the K-major SWIZZLE_128B layout of the GEMM's A operand, so the rest is gemm.cu's pipeline (1-CTA
// flavour: tile 128 x N, N = min(256, Cout) chosen at run time, 4-stage ring, two TMEM accumulator stages,
// warp-uniform MMA issue).
//
// Sub-pixel form of "nearest x2 upsample then conv (1,3,3) pad 1" (CosmosUpsample3d): output pixels of parity (py, px)
// are a 2x2 convolution of the *source-resolution* tensor with pre-summed weights, so the kernel iterates source
// positions (h, w) and stores to (2h + py, 2w + px) (`out_scale` = 2): 16 tap-GEMMs on quarter-size tiles instead of
// 9 on full size, and the 4x larger upsampled tensor is never written.
//
// Epilogue: + bias, optional residual (same position / frame-mapped / 2x2 or 2-frame average pool / nearest-upsampled
// source = the "+ avg_pool" and "+ x" terms of the resamplers and the resnet / attention skip), bf16 store, and
// per-frame sum / sum-of-squares of the stored values accumulated for the per-frame GroupNorm that follows.
//
// Roofline: tensor pipe, 2 * T*H*W * Cout * taps*Cin flop per launch.
#include <stdlib.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace drb {
namespace {

constexpr int kTileH = 8, kTileW = 16;          // 128 output positions per M tile
constexpr int kMaxN = 256, kBlockK = 64, kUmmaK = 16;
constexpr int kABytes = 128 * kBlockK * 2;      // 16 KB
// kCta = 1: one CTA per 128-position tile, W tile of up to 256 rows staged per CTA (48 KB stages, 4 of them).
// kCta = 2: a CTA pair (cta_group::2 MMA, M = 256 = two position tiles) shares the W tile — each CTA stages half of its rows
// (32 KB stages, 6 of them): a third less operand traffic per MMA, which is what lifted the (1,3,3) convolutions from
// 1370 - 1460 to 1530 - 1590 TFLOP/s (tensor pipe 96.8 % active).
// kSplit (pair flavour only, chosen for convolutions with few taps): the epilogue is TWO stages on different warps — four
// "drain" warps move the accumulator (+ bias, rounded to bf16) from tensor memory into a ring of four 16 KB buffers in
// shared memory (one buffer = 128 positions x 64 channels, a quarter of the widest tile), eight "finish" warps stream the
// buffers out (skip term, GroupNorm sums, fully coalesced 16-byte loads and stores) — because with 1 - 6 k MMA cycles per
// tile the single-stage epilogue (10 - 13 k cycles of a latency-bound TMEM -> registers -> shared -> global chain on two
// warps per scheduler) bounded those kernels.  The quarter-tile ring leaves room for a 5-stage operand ring: with the two
// whole-tile buffers of the first version only 3 stages fitted, and 96 KB of operands in flight per SM cannot cover the
// L2 latency at the MMA's consumption rate (ncu: the MMA warp waited for operands 37 % of the time, 7 us per tile against
// 3.6 us of MMA work).
template <int kCta, bool kSplit>
struct ConvCfg {
  static constexpr int kBBytes = kMaxN / kCta * kBlockK * 2;   // 32 KB / 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;        // 48 KB / 32 KB
  static constexpr int kStages = kSplit ? 5 : (kCta == 1 ? 4 : 6);
  static constexpr int kThreads = kSplit ? 512 : 384;
};
constexpr int kEpiWarps = 8;                    // single-stage epilogue: two per TMEM lane quadrant, each taking half of the tile's chunks
constexpr int kEpiStage = kEpiWarps * 4096;     // per epilogue warp: 2 x (32 pixels x 64 B = one 32-channel chunk), XOR-swizzled:
                                                // [0, 2048) output rows on their way out, [2048, 4096) skip-term rows on their way in
constexpr int kQuarterBytes = 128 * 128;        // split epilogue: 128 positions x 64 channels bf16, 16-byte slots XOR-swizzled by row
constexpr int kQuarters = 4;                    // ring of quarter-tile buffers between the drain and the finish warps
constexpr int kConvSmem = 4 * (kABytes + kMaxN * kBlockK * 2) + 1024 + 256 + kEpiStage;   // the same for all flavours
static_assert(5 * (kABytes + kMaxN / 2 * kBlockK * 2) + 256 + kQuarters * kQuarterBytes + 1024 <= kConvSmem, "split flavour must fit");

struct ConvMaps {
  CUtensorMap x[4];   // [0] for stride 1; [py*2 + px] parity views for stride 2
  CUtensorMap w;
};

struct ConvParams {
  int T_out, H_out, W_out;      // positions iterated (tile space)
  int out_H, out_W;             // dims of the output tensor (= H_out*out_scale when out_scale == 2)
  int Cin, Cout, block_n;
  int kt, kh, kw, pad_h, pad_w, stride_hw, tmode;
  int out_scale, out_off_h, out_off_w;
  __nv_bfloat16* out;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* resid;
  int resid_mode;
  int rH, rW;                   // dims of the residual source tensor (channels = Cout)
  double* stats;                // [T_out][2] (sum, sum of squares) or null
};

// source frame of output frame t for temporal tap dt (kt taps)
__device__ __forceinline__ int src_frame(int tmode, int t, int dt, int kt) {
  if (tmode == DRB_TMODE_DOWN2) return max(2 * t + dt - 2, 0);               // cat[x0, x], causal pad 1, stride 2
  if (tmode == DRB_TMODE_UP2) return (max(t + dt - (kt - 1), 0) + 1) >> 1;   // source is repeat_interleave(x, 2)[1:]
  return max(t + dt - (kt - 1), 0);                                          // causal: first frame replicated in front
}

template <int kCta, bool kSplit>
__global__ void __launch_bounds__((ConvCfg<kCta, kSplit>::kThreads), 1)
conv3d_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p) {
  constexpr int kStages = ConvCfg<kCta, kSplit>::kStages, kStageBytes = ConvCfg<kCta, kSplit>::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* q_full = tmem_empty_bar + 2;       // split epilogue: quarter buffer b holds drained columns / is free again
  uint64_t* q_empty = q_full + kQuarters;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + kQuarters);
  uint8_t* epi_stage = smem + kStages * kStageBytes + 256;   // single-stage epilogue staging, or the quarter-tile ring

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_w = (p.W_out + kTileW - 1) / kTileW, tiles_h = (p.H_out + kTileH - 1) / kTileH;
  const int tiles_n = (p.Cout + p.block_n - 1) / p.block_n;
  const int tiles_m = p.T_out * tiles_h * tiles_w;
  const int num_tiles = ((tiles_m + kCta - 1) / kCta) * tiles_n;   // units of work: one position tile (pair: two) x one n-tile
  const uint32_t cta_rank = kCta == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int unit_id = blockIdx.x / kCta, num_units = gridDim.x / kCta;
  const int cchunks = p.Cin / kBlockK;
  const int taps = p.kt * p.kh * p.kw;
  const int num_kb = taps * cchunks;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&maps.x[0]);
    prefetch_tmap(&maps.w);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], (kSplit ? 4 : kEpiWarps) * kCta);   // the TMEM-reading warps of every CTA of the group
    }
    for (int i = 0; i < kQuarters; ++i) {
      mbar_init(&q_full[i], 4);                                         // the four drain warps
      mbar_init(&q_empty[i], 8);                                        // the eight finish warps
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc<kCta>(tmem_slot, 512);
    tmem_relinquish<kCta>();
  }
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Tile order.  Spatial kernels: n-tile fastest, then w, h, t — the CTAs working on one spatial tile at the same time share
  // its activation boxes in L2, and neighbouring tiles share their halo.  Temporal kernels (kt > 1): the FRAME runs right
  // after the n-tile, so the CTAs running concurrently hold the same spatial tile at consecutive frames and the kt source
  // frames each of them reads are the ones its neighbours read too: every input box comes from HBM once (the spatial
  // order re-read it kt times: a frame is 29 - 58 MB, and input + output + skip of one output frame overflow the L2).
  // Returns false when this CTA's half of a pair unit lies past the last position tile (odd tile count): it then works on
  // a copy of the last tile and stores nothing.
  auto decode_tile = [&](int tile, int& t, int& h0, int& w0, int& n0) -> bool {
    const int tn = tile % tiles_n;
    int m = (tile / tiles_n) * kCta + static_cast<int>(cta_rank);
    const bool valid = m < tiles_m;
    if (!valid) m = tiles_m - 1;
    if (p.kt > 1) {
      t = m % p.T_out;
      m /= p.T_out;
    }
    const int tw = m % tiles_w;
    m /= tiles_w;
    const int th = m % tiles_h;
    if (p.kt <= 1) t = m / tiles_h;
    h0 = th * kTileH;
    w0 = tw * kTileW;
    n0 = tn * p.block_n;
    return valid;
  };

  // The eight epilogue warps are what bounds the few-tap convolutions, and at the 168 registers a 384-thread CTA allows
  // the compiler spilled the TMEM address and re-derived shared-memory addresses from SR_TID in every chunk (ncu:
  // long-scoreboard stalls on LDL / S2R).  Warps 0-3 (one TMA thread, one MMA thread, two idle) hand their registers over.
  if (warp_idx < 4) {
  reg_dec<kSplit ? 56 : 72>();
  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The WHOLE warp runs the loop in uniform control flow and one elected lane issues (as the MMA warp does): box
    // coordinates, barrier and shared-memory addresses then live in uniform registers, which is what UTMALDG reads.  A
    // lane-divergent producer paid an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence per operand — about 60 single-lane
    // instructions per k-block, which on a scheduler shared with busy drain and finish warps took 880 cycles against the
    // 512 the k-block's MMAs need (ncu, split flavour: the producer warp was issuing in 80 % of its samples).
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      const int b_rows = p.block_n / kCta;                               // W rows this CTA stages
      const uint32_t stage_tx = kABytes + b_rows * kBlockK * 2;
      for (int tile = unit_id; tile < num_tiles; tile += num_units) {
        int t, h0, w0, n0;
        decode_tile(tile, t, h0, w0, n0);
        for (int tap = 0; tap < taps; ++tap) {
          const int dx = tap % p.kw, dy = (tap / p.kw) % p.kh, dt = tap / (p.kw * p.kh);
          const int ts = src_frame(p.tmode, t, dt, p.kt);
          int hs, wsrc, mi = 0;
          if (p.stride_hw == 1) {
            hs = h0 + dy - p.pad_h;
            wsrc = w0 + dx - p.pad_w;
          } else {   // source row 2h + r, r = dy - pad: parity view (r & 1), row h + (r - parity) / 2
            const int ry = dy - p.pad_h, rx = dx - p.pad_w;
            const int py = ry & 1, px = rx & 1;
            hs = h0 + ((ry - py) >> 1);
            wsrc = w0 + ((rx - px) >> 1);
            mi = py * 2 + px;
          }
          for (int cc = 0; cc < cchunks; ++cc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kStageBytes;
            if (issuer) {
              if constexpr (kCta == 1) {
                mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                tma_load_4d(sa, &maps.x[mi], &full_bar[stage], cc * kBlockK, wsrc, hs, ts);
                tma_load_2d(sa + kABytes, &maps.w, &full_bar[stage], tap * p.Cin + cc * kBlockK, n0);
              } else {   // both CTAs' bytes are accounted on the leader's barrier
                if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], stage_tx * 2);
                tma_load_4d_pair(sa, &maps.x[mi], &full_bar[stage], cc * kBlockK, wsrc, hs, ts);
                tma_load_2d_pair(sa + kABytes, &maps.w, &full_bar[stage], tap * p.Cin + cc * kBlockK,
                                 n0 + static_cast<int>(cta_rank) * b_rows);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane; leader CTA only)
    if (is_leader) {
    const uint32_t idesc = make_idesc_bf16(128 * kCta, p.block_n);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const bool issuer = elect_one();
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_lo = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);
    int stage = 0, iter = 0;
    uint32_t phase = 0;
    for (int tile = unit_id; tile < num_tiles; tile += num_units, ++iter) {
      const int as = iter & 1;
      mbar_wait(&tmem_empty_bar[as], ((iter >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tm + as * kMaxN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = smem_lo + stage * (kStageBytes >> 4);
        const uint32_t b_lo = a_lo + (kABytes >> 4);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_ss<kCta>(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k),
                          (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k), idesc, (kb | k) != 0);
          if constexpr (kCta == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_pair(&empty_bar[stage], 0x3);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (issuer) {
        if constexpr (kCta == 1) umma_commit(&tmem_full_bar[as]);
        else umma_commit_pair(&tmem_full_bar[as], 0x3);
      }
      __syncwarp();
    }
    }
  }
  } else if (kSplit && warp_idx < 8) {
    // ------------------------------------------------------------------ split epilogue, stage 1: drain (one warp per TMEM lane quadrant)
    reg_inc<160>();
    const int q = warp_idx & 3;
    const int r = q * 32 + lane;
    const int nch = (p.block_n + 31) / 32, nq = (nch + 1) >> 1;
    const uint32_t ring_u32 = smem_u32(epi_stage);
    const uint32_t row_off = r * 128, swz = r & 7;
    int iter = 0;
    uint32_t qc = 0;       // quarters handed over so far: buffer qc % 4, phase (qc / 4) & 1
    for (int tile = unit_id; tile < num_tiles; tile += num_units, ++iter) {
      int t, h0, w0, n0;
      decode_tile(tile, t, h0, w0, n0);
      const int as = iter & 1;
      mbar_wait(&tmem_full_bar[as], (iter >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kMaxN;
      uint32_t accbuf[2][32];
      tmem_ld32(taddr, accbuf[0]);
#pragma unroll
      for (int c = 0; c < kMaxN / 32; ++c) {
        if (c >= nch) break;   // warp-uniform
        uint4 bias4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cg = n0 + c * 32 + g * 8;
          bias4[g] = cg < p.Cout ? __ldg(reinterpret_cast<const uint4*>(p.bias + cg)) : make_uint4(0u, 0u, 0u, 0u);
        }
        const uint32_t qi = qc + (c >> 1);
        if ((c & 1) == 0) mbar_wait(&q_empty[qi & 3], ((qi >> 2) & 1) ^ 1);
        tmem_wait_ld();
        uint32_t (&acc)[32] = accbuf[c & 1];
        if (c + 1 < nch) {
          tmem_ld32(taddr + (c + 1) * 32, accbuf[(c + 1) & 1]);
        } else {               // the accumulator stage is free as soon as its last columns are in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCta == 1) mbar_arrive(&tmem_empty_bar[as]);
            else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
          }
        }
        const uint32_t dst = ring_u32 + (qi & 3) * kQuarterBytes + row_off;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t bb[4] = {bias4[g].x, bias4[g].y, bias4[g].z, bias4[g].w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = pack_bf16x2(__uint_as_float(acc[g * 8 + 2 * j]) + bf16_lo(bb[j]), __uint_as_float(acc[g * 8 + 2 * j + 1]) + bf16_hi(bb[j]));
          st_shared_v4(dst + ((((c & 1) * 4 + g) ^ swz) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        }
        if ((c & 1) == 1 || c + 1 == nch) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&q_full[qi & 3]);
        }
      }
      qc += nq;
    }
  } else if (kSplit) {
    // ------------------------------------------------------------------ split epilogue, stage 2: finish (eight warps, streaming)
    // Per quarter buffer (128 positions x 64 channels) thread f takes the 16-byte channel slot f % 8 of the positions
    // f / 8 + 32 u, u = 0..3: a warp reads and writes whole 128-byte lines — skip term in, output out.  The positions are
    // the same for every quarter of a tile, so everything per vector is one 32-bit add on offsets computed once per tile.
    // This stage bounded the few-tap convolutions; ncu (r02) showed its first version issue-bound at 200 instructions per
    // vector (64-bit address arithmetic, a runtime division and five range checks per vector) and, once unrolled, missing
    // the instruction cache.  The skip vectors of the NEXT quarter — across the tile boundary too — are requested before
    // the current one is processed.
    reg_inc<136>();
    const int f = threadIdx.x - 256;
    const int slot = f & 7, prow = f >> 3;                        // rows prow + 32 u: (row & 7) is the same for all four
    const int nq = (p.block_n + 63) >> 6;
    const int rmode = p.resid_mode;
    const bool single = rmode == DRB_RES_SAME || rmode == DRB_RES_FRAME_UP2 || rmode == DRB_RES_NEAREST_UP_HW;
    const int rsh = rmode == DRB_RES_NEAREST_UP_HW ? 1 : 0;
    const int ssh = p.out_scale - 1;                              // out_scale is 1 or 2
    const int64_t out_frame = static_cast<int64_t>(p.out_H) * p.out_W * p.Cout, res_frame = static_cast<int64_t>(p.rH) * p.rW * p.Cout;
    const uint32_t my_u32 = smem_u32(epi_stage) + prow * 128 + ((slot ^ (prow & 7)) << 4);

    struct TileCtx {
      int t, h0, w0, n0;
      bool valid;
      int rel_o[4];                    // element offset of position u in its output frame, -1: nothing to store
      int rel_r[4];                    // the same in the skip-term frame (single-source skip terms)
      __nv_bfloat16* out_t;            // output frame + n0 + this thread's channel slot
      const __nv_bfloat16* res_t;      // skip-term frame + n0 + slot
    };
    auto make_ctx = [&](int tile) {
      TileCtx c;
      c.valid = decode_tile(tile, c.t, c.h0, c.w0, c.n0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int row = prow + 32 * u;
        const int h = c.h0 + (row >> 4), w = c.w0 + (row & 15);
        const int oh = h * p.out_scale + p.out_off_h, ow = w * p.out_scale + p.out_off_w;
        c.rel_o[u] = (c.valid && h < p.H_out && w < p.W_out) ? (oh * p.out_W + ow) * p.Cout : -1;
        c.rel_r[u] = ((oh >> rsh) * p.rW + (ow >> rsh)) * p.Cout;
      }
      c.out_t = p.out + static_cast<int64_t>(c.t) * out_frame + c.n0 + slot * 8;
      c.res_t = single ? p.resid + static_cast<int64_t>(rmode == DRB_RES_FRAME_UP2 ? (c.t + 1) >> 1 : c.t) * res_frame + c.n0 + slot * 8
                       : nullptr;
      return c;
    };
    // channels [ch, ch + 8) of quarter qq exist in this n-tile
    auto col_ok = [&](const TileCtx& c, int qq) { const int ch = qq * 64 + slot * 8; return ch < p.block_n && c.n0 + ch < p.Cout; };
    uint4 pf[4] = {};        // skip vectors of the quarter processed next
    auto request_skips = [&](const TileCtx& c, int qq) {
      const bool cok = col_ok(c, qq);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        pf[u] = (cok && c.rel_o[u] >= 0) ? __ldg(reinterpret_cast<const uint4*>(c.res_t + c.rel_r[u] + qq * 64)) : make_uint4(0u, 0u, 0u, 0u);
    };

    int tile = unit_id;
    uint32_t qc = 0;
    TileCtx c{};
    if (tile < num_tiles) {
      c = make_ctx(tile);
      if (single) request_skips(c, 0);
    }
    while (tile < num_tiles) {
      const int next_tile = tile + num_units;
      TileCtx nx = c;
      uint64_t s1p = f32x2_pack(0.f, 0.f), s2p = s1p;    // GroupNorm sums of the even / odd channels (packed fp32x2 lanes)
#pragma unroll 1
      for (int qq = 0; qq < nq; ++qq, ++qc) {
        const uint32_t src = my_u32 + (qc & 3) * kQuarterBytes;
        uint4 cur[4], xv[4];
        mbar_wait(&q_full[qc & 3], (qc >> 2) & 1);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          cur[u] = pf[u];
          xv[u] = ld_shared_v4(src + u * 4096);
        }
        if (qq + 1 < nq) {
          if (single) request_skips(c, qq + 1);
        } else if (next_tile < num_tiles) {
          nx = make_ctx(next_tile);
          if (single) request_skips(nx, 0);
        }
        const bool cok = col_ok(c, qq);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!cok || c.rel_o[u] < 0) continue;
          uint32_t o[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
          if (single) {
            // the drained value is the convolution output already rounded to bf16, as the reference has it before the add
            const uint32_t rr[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = bf16x2_add(o[j], rr[j]);   // == bf16(fp32(a) + fp32(b)): the fp32 sum of two bf16 is
                                                                          // exact or rounds to the larger operand, a bf16 already
          } else if (rmode != DRB_RES_NONE) {       // pooled skip terms (two resampling convolutions per net): 2 or 4 rows averaged
            const int row = prow + 32 * u;
            const int oh = (c.h0 + (row >> 4)) * p.out_scale + p.out_off_h, ow = (c.w0 + (row & 15)) * p.out_scale + p.out_off_w;
            const int nsrc = rmode == DRB_RES_POOL_HW ? 4 : 2;
            const float rscale = rmode == DRB_RES_POOL_HW ? 0.25f : 0.5f;
            float ra[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int i2 = 0; i2 < nsrc; ++i2) {
              int tt = c.t, hh = oh, ww = ow;
              if (rmode == DRB_RES_POOL_HW) {
                hh = 2 * oh + (i2 >> 1);
                ww = 2 * ow + (i2 & 1);
                if (hh >= p.rH || ww >= p.rW) continue;
              } else {
                tt = i2 == 0 ? max(2 * c.t - 1, 0) : 2 * c.t;
              }
              const uint4 rv = *reinterpret_cast<const uint4*>(p.resid + ((static_cast<int64_t>(tt) * p.rH + hh) * p.rW + ww) * p.Cout +
                                                               c.n0 + qq * 64 + slot * 8);
              const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                ra[2 * j] += bf16_lo(rr[j]);
                ra[2 * j + 1] += bf16_hi(rr[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = pack_bf16x2(bf16_lo(o[j]) + bf16_round(ra[2 * j] * rscale), bf16_hi(o[j]) + bf16_round(ra[2 * j + 1] * rscale));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ab = f32x2_pack(bf16_lo(o[j]), bf16_hi(o[j]));
            s1p = f32x2_add(s1p, ab);
            s2p = f32x2_fma(ab, ab, s2p);
          }
          *reinterpret_cast<uint4*>(c.out_t + c.rel_o[u] + qq * 64) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_empty[qc & 3]);
      }
      if (p.stats != nullptr && c.valid) {
        float s1, s2, s1b, s2b;
        f32x2_unpack(s1p, s1, s1b);
        f32x2_unpack(s2p, s2, s2b);
        s1 += s1b;
        s2 += s2b;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
          atomicAdd(&p.stats[2 * c.t], static_cast<double>(s1));
          atomicAdd(&p.stats[2 * c.t + 1], static_cast<double>(s2));
        }
      }
      tile = next_tile;
      c = nx;
    }
  } else {
    reg_inc<208>();
    // ------------------------------------------------------------------ epilogue
    // Eight warps: warp w reads TMEM lane quadrant w % 4 (its 32 output positions) and takes half of the tile's 32-column
    // chunks.  With few taps (1x1x1, 3x1x1) a tile's MMAs take 1 - 6 k cycles and the epilogue — bias, skip term, bf16
    // store, GroupNorm sums — is what bounds the kernel, so (a) it is spread over twice the warps and (b) the skip-term
    // rows are fetched into registers BEFORE the wait for the accumulator, i.e. under the tile's MMAs.
    const int q = warp_idx & 3;
    const int ew = warp_idx - 4;
    const int nch = (p.block_n + 31) / 32, per = (nch + 1) / 2;
    const int c_begin = (ew >> 2) * per, c_end = min(nch, c_begin + per);
    uint8_t* my_stage = epi_stage + ew * 4096;
    uint8_t* res_stage = my_stage + 2048;
    int iter = 0;
    for (int tile = unit_id; tile < num_tiles; tile += num_units, ++iter) {
      int t, h0, w0, n0;
      const bool tile_valid = decode_tile(tile, t, h0, w0, n0);
      const int as = iter & 1;
      const int r = q * 32 + lane;
      const int h = h0 + r / kTileW, w = w0 + r % kTileW;
      const bool ok = tile_valid && h < p.H_out && w < p.W_out;
      const int oh = h * p.out_scale + p.out_off_h, ow = w * p.out_scale + p.out_off_w;
      // residual source rows (up to 4 averaged)
      const __nv_bfloat16* rrow[4] = {nullptr, nullptr, nullptr, nullptr};
      int nres = 0;
      float rscale = 1.0f;
      if (ok) {
        auto rpix = [&](int tt, int hh, int ww) {
          return p.resid + ((static_cast<int64_t>(tt) * p.rH + hh) * p.rW + ww) * p.Cout;
        };
        if (p.resid_mode == DRB_RES_SAME) {
          rrow[0] = rpix(t, oh, ow);
          nres = 1;
        } else if (p.resid_mode == DRB_RES_FRAME_UP2) {      // x'[t] = x[(t + 1) / 2]
          rrow[0] = rpix((t + 1) >> 1, oh, ow);
          nres = 1;
        } else if (p.resid_mode == DRB_RES_NEAREST_UP_HW) {  // x''[oh, ow] = x'[oh / 2, ow / 2]
          rrow[0] = rpix(t, oh >> 1, ow >> 1);
          nres = 1;
        } else if (p.resid_mode == DRB_RES_POOL_HW) {        // 2x2 average of the zero-padded source
          nres = 4;
          rscale = 0.25f;
          for (int i = 0; i < 4; ++i) {
            const int hh = 2 * oh + (i >> 1), ww = 2 * ow + (i & 1);
            rrow[i] = (hh < p.rH && ww < p.rW) ? rpix(t, hh, ww) : nullptr;
          }
        } else if (p.resid_mode == DRB_RES_POOL_T) {         // frames max(2t-1, 0) and 2t of the source (cat[x0, x] pooled)
          nres = 2;
          rscale = 0.5f;
          rrow[0] = rpix(max(2 * t - 1, 0), oh, ow);
          rrow[1] = rpix(2 * t, oh, ow);
        }
      }
      // Single-source skip terms (resnet / attention skip, the "+ x" of the upsamplers) are fetched COALESCED: per 32-channel
      // chunk the warp reads its 32 pixels x 64 B with 4 lanes per pixel (8 pixels = 8 half-lines per instruction, instead of
      // 32 lanes on 32 different lines — that form saturated the LSU: ncu lg_throttle, +75 % on the (3,1,1) convolutions),
      // parks them in shared memory and every thread picks up its own row.  The loads of chunk c + 1 are issued before
      // chunk c is processed; those of the first chunk before the wait for the accumulator.
      const bool pre = p.resid_mode == DRB_RES_SAME || p.resid_mode == DRB_RES_FRAME_UP2 || p.resid_mode == DRB_RES_NEAREST_UP_HW;
      const __nv_bfloat16* rp[4] = {nullptr, nullptr, nullptr, nullptr};
      if (pre) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r2 = q * 32 + it * 8 + (lane >> 2);
          const int h2 = h0 + r2 / kTileW, w2 = w0 + r2 % kTileW;
          if (tile_valid && h2 < p.H_out && w2 < p.W_out) {
            int tt = t, hh = h2 * p.out_scale + p.out_off_h, ww = w2 * p.out_scale + p.out_off_w;
            if (p.resid_mode == DRB_RES_FRAME_UP2) tt = (t + 1) >> 1;
            if (p.resid_mode == DRB_RES_NEAREST_UP_HW) { hh >>= 1; ww >>= 1; }
            rp[it] = p.resid + ((static_cast<int64_t>(tt) * p.rH + hh) * p.rW + ww) * p.Cout + n0 + (lane & 3) * 8;
          }
        }
      }
      uint4 rnext[4];
      auto issue_skip = [&](int c) {
        const int cl = c * 32 + (lane & 3) * 8;
        const bool live = c < c_end && cl < p.block_n && n0 + cl < p.Cout;
#pragma unroll
        for (int it = 0; it < 4; ++it)
          rnext[it] = (live && rp[it] != nullptr) ? *reinterpret_cast<const uint4*>(rp[it] + c * 32) : make_uint4(0u, 0u, 0u, 0u);
      };
      if (pre) issue_skip(c_begin);
      mbar_wait(&tmem_full_bar[as], (iter >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kMaxN;
      float s1 = 0.f, s2 = 0.f;
      // the chunks are software-pipelined: while chunk c is processed, the accumulator columns of chunk c + 1 are already on
      // their way out of tensor memory (two register buffers, statically indexed through the full unroll)
      uint32_t accbuf[2][32];
      if (c_begin < c_end && n0 + c_begin * 32 < p.Cout) tmem_ld32(taddr + c_begin * 32, accbuf[0]);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = c_begin + ci;
        const int col = n0 + c * 32;
        if (c >= c_end || col >= p.Cout) break;   // warp-uniform
        uint4 rmine[4];
        if (pre) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int pr = it * 8 + (lane >> 2);
            *reinterpret_cast<uint4*>(res_stage + pr * 64 + (((lane & 3) ^ ((pr >> 1) & 3)) * 16)) = rnext[it];
          }
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g)
            rmine[g] = *reinterpret_cast<const uint4*>(res_stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) * 16));
          issue_skip(c + 1);              // in flight while this chunk is processed
        }
        // bias of this chunk: requested before the wait for the accumulator columns
        uint4 bias4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          bias4[g] = (col + g * 8 < p.Cout) ? __ldg(reinterpret_cast<const uint4*>(p.bias + col + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
        tmem_wait_ld();
        uint32_t (&acc)[32] = accbuf[ci & 1];
        if (c + 1 < c_end && col + 32 < p.Cout) tmem_ld32(taddr + (c + 1) * 32, accbuf[(ci + 1) & 1]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          // Straight-line on purpose: lanes without a pixel (ragged tiles) and column groups past Cout compute on finite
          // dummies and are masked out of the GroupNorm sums / never stored, so the four groups of a chunk interleave
          // instead of being four divergence-guarded blocks (the epilogue ran at 6.8 cycles per instruction, ncu).
          const int cg = col + g * 8;
          const float live = (ok && cg < p.Cout && c * 32 + g * 8 < p.block_n) ? 1.0f : 0.0f;
          float v[8];
          const uint4 bv = bias4[g];
          const uint32_t bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[2 * j] = __uint_as_float(acc[g * 8 + 2 * j]) + bf16_lo(bb[j]);
            v[2 * j + 1] = __uint_as_float(acc[g * 8 + 2 * j + 1]) + bf16_hi(bb[j]);
          }
          if (p.resid_mode != DRB_RES_NONE) {       // uniform
            float ra[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (pre) {
              const uint4 rv = rmine[g];
              const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                ra[2 * j] = bf16_lo(rr[j]);
                ra[2 * j + 1] = bf16_hi(rr[j]);
              }
            } else if (live != 0.0f) {              // pooled skip terms (two resampling convolutions per net): per-lane sources
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (i >= nres || rrow[i] == nullptr) continue;
                const uint4 rv = *reinterpret_cast<const uint4*>(rrow[i] + cg);
                const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  ra[2 * j] += bf16_lo(rr[j]);
                  ra[2 * j + 1] += bf16_hi(rr[j]);
                }
              }
            }
            // the reference rounds the convolution output to bf16 before adding the skip term
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = bf16_round(v[j]) + bf16_round(ra[j] * rscale);
          }
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            const float a = bf16_lo(o[j]), b = bf16_hi(o[j]);
            s1 = fmaf(live, a + b, s1);
            s2 = fmaf(live * a, a, fmaf(live * b, b, s2));
          }
          *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        // this warp's 32 pixels are two rows of 16 consecutive pixels: store 64-byte channel segments, 8 pixels per instruction
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + (lane >> 2), ch = lane & 3;
          const int ph = h0 + (q * 32 + rr) / kTileW, pw = w0 + (q * 32 + rr) % kTileW;
          if (tile_valid && ph < p.H_out && pw < p.W_out && col + ch * 8 < p.Cout && c * 32 + ch * 8 < p.block_n) {
            const int64_t px = (static_cast<int64_t>(t) * p.out_H + ph * p.out_scale + p.out_off_h) * p.out_W + pw * p.out_scale + p.out_off_w;
            *reinterpret_cast<uint4*>(p.out + px * p.Cout + col + ch * 8) =
                *reinterpret_cast<const uint4*>(my_stage + rr * 64 + ((ch ^ ((rr >> 1) & 3)) * 16));
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCta == 1) mbar_arrive(&tmem_empty_bar[as]);
        else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      }
      if (p.stats != nullptr && tile_valid) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
          atomicAdd(&p.stats[2 * t], static_cast<double>(s1));
          atomicAdd(&p.stats[2 * t + 1], static_cast<double>(s2));
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync(); else __syncthreads();   // no CTA leaves while its peer can still signal into it
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<kCta>(tmem_base, 512);
  }
}

template <int kCta, bool kSplit>
int launch_conv(const ConvMaps& maps, const ConvParams& p, int units, cudaStream_t stream) {
  auto kernel = conv3d_kernel<kCta, kSplit>;
  static DeviceOnce configured;   // per flavour and per device: the attribute belongs to the device's context
  const int rc = device_once(configured, [&] {
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem), "conv3d_kernel smem");
  });
  if (rc) return rc;
  int groups = num_sms() / kCta;
  if (groups > units) groups = units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * kCta);
  cfg.blockDim = dim3(ConvCfg<kCta, kSplit>::kThreads);
  cfg.dynamicSmemBytes = kConvSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DRB_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, p));
  return 0;
}

}  // namespace
}  // namespace drb

extern "C" int drb_conv3d_cl(const drb_conv3d_args* a, void* stream) {
  using namespace drb;
  DRB_REQUIRE(a != nullptr, "null argument block");
  DRB_REQUIRE(a->x && a->w && a->bias && a->out, "null pointer");
  DRB_REQUIRE(a->T_in > 0 && a->H_in > 0 && a->W_in > 0 && a->T_out > 0 && a->H_out > 0 && a->W_out > 0, "bad tensor dims");
  DRB_REQUIRE(a->Cin > 0 && a->Cin % 64 == 0, "Cin must be a multiple of 64 (zero-pad narrower tensors)");
  DRB_REQUIRE(a->Cout > 0 && a->Cout % 16 == 0, "Cout must be a multiple of 16");
  DRB_REQUIRE(a->kt >= 1 && a->kt <= 3 && a->kh >= 1 && a->kh <= 3 && a->kw >= 1 && a->kw <= 3, "taps must be 1..3 per axis");
  DRB_REQUIRE(a->stride_hw == 1 || a->stride_hw == 2, "spatial stride must be 1 or 2");
  DRB_REQUIRE(a->pad_h >= 0 && a->pad_h <= 2 && a->pad_w >= 0 && a->pad_w <= 2, "bad padding");
  DRB_REQUIRE(a->tmode >= DRB_TMODE_CAUSAL && a->tmode <= DRB_TMODE_UP2, "unknown tmode");
  DRB_REQUIRE(a->resid_mode >= DRB_RES_NONE && a->resid_mode <= DRB_RES_NEAREST_UP_HW, "unknown resid_mode");
  DRB_REQUIRE((a->resid_mode == DRB_RES_NONE) == (a->resid == nullptr), "resid pointer and resid_mode disagree");
  DRB_REQUIRE(a->out_scale == 1 || a->out_scale == 2, "out_scale must be 1 or 2");
  DRB_REQUIRE(a->out_off_h >= 0 && a->out_off_h < a->out_scale && a->out_off_w >= 0 && a->out_off_w < a->out_scale, "bad output offset");
  DRB_REQUIRE(a->out_scale == 1 || a->stride_hw == 1, "out_scale 2 requires stride 1");
  DRB_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->resid) & 15) == 0,
              "out / bias / resid must be 16-byte aligned");
  // temporal extent the frame mapping may touch
  {
    const int t_last = a->T_out - 1;
    int need;
    if (a->tmode == DRB_TMODE_DOWN2) need = (a->kt == 1) ? 2 * t_last - 2 : 2 * t_last;
    else if (a->tmode == DRB_TMODE_UP2) need = (t_last + 1) >> 1;
    else need = t_last;
    if (need < 0) need = 0;
    DRB_REQUIRE(need < a->T_in, "temporal mapping reads past the last source frame");
  }
  const uint64_t C = static_cast<uint64_t>(a->Cin), W = static_cast<uint64_t>(a->W_in), H = static_cast<uint64_t>(a->H_in),
                 T = static_cast<uint64_t>(a->T_in);
  ConvMaps maps;
  int rc;
  const uint32_t box[4] = {kBlockK, kTileW, kTileH, 1};
  if (a->stride_hw == 1) {
    const uint64_t dims[4] = {C, W, H, T};
    const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    rc = make_tmap_nd_bf16(&maps.x[0], a->x, 4, dims, strides, box);
    if (rc) return rc;
    for (int i = 1; i < 4; ++i) maps.x[i] = maps.x[0];
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const uint64_t wp = (W + 1 - px) / 2, hp = (H + 1 - py) / 2;   // ceil((W - px) / 2)
        CUtensorMap* m = &maps.x[py * 2 + px];
        if (wp == 0 || hp == 0) {       // no source pixel of this parity (H or W == 1): never selected by a valid tap
          *m = maps.x[0];
          continue;
        }
        const uint64_t dims[4] = {C, wp, hp, T};
        const uint64_t strides[3] = {2 * C * 2, 2 * W * C * 2, H * W * C * 2};
        const char* base = static_cast<const char*>(a->x) + (static_cast<uint64_t>(py) * W + px) * C * 2;
        rc = make_tmap_nd_bf16(m, base, 4, dims, strides, box);
        if (rc) return rc;
      }
  }
  const int taps = a->kt * a->kh * a->kw;
  const int block_n = a->Cout <= kMaxN ? a->Cout : kMaxN;
  const int tiles_m = a->T_out * ((a->H_out + kTileH - 1) / kTileH) * ((a->W_out + kTileW - 1) / kTileW);
  const int tiles_n = (a->Cout + block_n - 1) / block_n;
  // the CTA-pair flavour needs two position tiles to pair and a W tile that splits into two UMMA-legal halves
  static const int pair_ok = [] { const char* e = getenv("DRB_CONV_PAIR"); return e ? atoi(e) : 1; }();   // 0: A/B measurements
  const int cta = (pair_ok && tiles_m >= 2 && block_n % 32 == 0) ? 2 : 1;
  rc = make_tmap_2d_bf16(&maps.w, a->w, a->Cout, static_cast<uint64_t>(taps) * a->Cin, static_cast<uint64_t>(taps) * a->Cin,
                         block_n / cta, kBlockK);
  if (rc) return rc;
  ConvParams p{};
  p.T_out = a->T_out; p.H_out = a->H_out; p.W_out = a->W_out;
  p.out_H = a->H_out * a->out_scale; p.out_W = a->W_out * a->out_scale;
  p.Cin = a->Cin; p.Cout = a->Cout; p.block_n = block_n;
  p.kt = a->kt; p.kh = a->kh; p.kw = a->kw; p.pad_h = a->pad_h; p.pad_w = a->pad_w; p.stride_hw = a->stride_hw; p.tmode = a->tmode;
  p.out_scale = a->out_scale; p.out_off_h = a->out_off_h; p.out_off_w = a->out_off_w;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.bias = static_cast<const __nv_bfloat16*>(a->bias);
  p.resid = static_cast<const __nv_bfloat16*>(a->resid);
  p.resid_mode = a->resid_mode;
  p.rH = a->resid_H; p.rW = a->resid_W;
  p.stats = a->stats;
  const int units = ((tiles_m + cta - 1) / cta) * tiles_n;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // The split, two-stage epilogue pays where the single-stage one bounds the kernel: up to 36 k-blocks per tile (1x1x1,
  // 3x1x1, the 2x2 sub-pixel convolutions, 3x3 at <= 256 channels: tools/conv_t_probe.py, tools/vae_kernel_probe.py, r02 —
  // (3,1,1) 256 -> 256 with skip 0.57 -> 0.35 ms, 512 -> 512 with skip 0.187 -> 0.147, (1,2,2) 512 -> 512 0.37 -> 0.32,
  // (1,3,3) 256 -> 256 -4 %; at 72 k-blocks the deeper operand ring of the single-stage flavour wins).  Pooled skip terms
  // (2 or 4 source rows per output vector) are slow in the streaming finish stage: those only up to 12 k-blocks.
  static const int split_ok = [] { const char* e = getenv("DRB_CONV_SPLIT"); return e ? atoi(e) : 1; }();   // 0: A/B measurements
  static const int split_kb = [] { const char* e = getenv("DRB_CONV_SPLIT_KB"); return e ? atoi(e) : 36; }();
  const int num_kb = taps * (a->Cin / kBlockK);
  const bool pooled = a->resid_mode == DRB_RES_POOL_HW || a->resid_mode == DRB_RES_POOL_T;
  // the finish stage addresses inside one frame with 32-bit element offsets
  const bool frames_fit = static_cast<int64_t>(p.out_H) * p.out_W * p.Cout < (int64_t(1) << 31) &&
                          static_cast<int64_t>(p.rH) * p.rW * p.Cout < (int64_t(1) << 31);
  const bool split = cta == 2 && split_ok && frames_fit && num_kb <= (pooled ? (split_kb < 12 ? split_kb : 12) : split_kb);
  if (split) return launch_conv<2, true>(maps, p, units, s);
  return cta == 2 ? launch_conv<2, false>(maps, p, units, s) : launch_conv<1, false>(maps, p, units, s);
}
