// Context parallelism over NVLink peer memory: one video's token sequence is split contiguously over P
// GPUs; every operator of the GeneralDIT is token-local except self-attention, which needs all keys of a head.  The
// Ulysses exchange (tokens-sharded <-> heads-sharded) is not a separate collective here: it is fused into the kernels on
// either side of the attention,
//
//   drb_gemm_qkv_norm_rope       (gemm.cu) the QKV projection whose epilogue normalises / rotates q and k and stores each head's
//                                q / k / v row straight into the peer that owns the head ([S, 3*D/P] buffer, P2P stores over
//                                NVLink) — the "all-to-all" is the GEMM's output;
//   drb_cp_qk_norm_rope_scatter  the same exchange as a stand-alone kernel after a plain QKV GEMM (CleanGeneralDIT.py:288-297,
//                                :45-84); kept as the A/B reference of the fused form (74.9 vs 73.6 ms per step at P = 8);
//   drb_attention_bf16_cp        (attention.cu) flash attention over the H/P local heads and all S tokens whose epilogue
//                                stores each output row into the peer that owns the token;
//   drb_cp_barrier               system-scope flag exchange between the P GPUs (one tiny kernel, no host involvement).
//
// Peer buffers are plain cudaMalloc allocations shared through CUDA IPC handles (one process per GPU); torch.distributed
// only carries the 64-byte handles at start-up.
#include <string.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace drb {
namespace {

struct PeerPtrs {
  void* p[DRB_CP_MAX_RANKS];
};

__device__ __forceinline__ float warp_sum_cp(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per local token row; 8 warps sweep the 3*H head slots (q heads, k heads, v heads).  Same arithmetic and
// rounding points as qk_norm_rope_kernel (elementwise.cu); only the destination differs.
__global__ void __launch_bounds__(256)
qk_norm_rope_scatter_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, const __nv_bfloat16* __restrict__ wq,
                            const __nv_bfloat16* __restrict__ wk, const __nv_bfloat16* __restrict__ cos_tab,
                            const __nv_bfloat16* __restrict__ sin_tab, int H, const PeerPtrs dst, int64_t dst_ld, int row0,
                            int heads_per_rank) {
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
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(row) * ld;
  const int64_t sect_w = static_cast<int64_t>(heads_per_rank) * 128;   // width of the q (k, v) section of a peer row
  for (int slot = warp; slot < 3 * H; slot += 8) {
    const int sect = slot / H, head = slot - sect * H;   // 0 q, 1 k, 2 v
    const uint2 xv = *reinterpret_cast<const uint2*>(base + static_cast<int64_t>(slot) * 128 + lane * 4);
    uint2 outv = xv;
    if (sect < 2) {
      const uint2 wv = sect == 1 ? wkv : wqv;
      const float w[4] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y)};
      float x[4] = {bf16_lo(xv.x), bf16_hi(xv.x), bf16_lo(xv.y), bf16_hi(xv.y)};
      const float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
      const float inv = rsqrtf(warp_sum_cp(ss) * (1.0f / 128.0f) + 1e-6f);
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
      outv = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
    }
    const int owner = head / heads_per_rank, hh = head - owner * heads_per_rank;
    __nv_bfloat16* drow = static_cast<__nv_bfloat16*>(dst.p[owner]) + static_cast<int64_t>(row0 + row) * dst_ld;
    *reinterpret_cast<uint2*>(drow + sect * sect_w + hh * 128 + lane * 4) = outv;
  }
}

// flags[r] of every peer <- epoch (release, system scope); then wait until all of my flags reach epoch (acquire).
__global__ void __launch_bounds__(32)
cp_barrier_kernel(const PeerPtrs flags, int rank, int world, uint32_t epoch) {
  const int j = threadIdx.x;
  if (j < world) {
    uint32_t* remote = static_cast<uint32_t*>(flags.p[j]) + rank;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const uint32_t* mine = static_cast<const uint32_t*>(flags.p[rank]) + j;
    const uint64_t t0 = globaltimer_ns();
    uint32_t v;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > 60000000000ull) {
        // a lost peer: report through the local status word (the host checks it) and carry on — a trap would kill the
        // CUDA context of the whole process
        static_cast<uint32_t*>(flags.p[rank])[DRB_CP_STATUS_WORD] = 1u;
        break;
      }
    } while (static_cast<int32_t>(v - epoch) < 0);
  }
}

}  // namespace
}  // namespace drb

using namespace drb;

extern "C" int drb_cp_qk_norm_rope_scatter(const void* qkv, int64_t ld, const void* wq, const void* wk, const void* cos_tab,
                                           const void* sin_tab, int S_local, int num_heads, void* const* dst_ptrs, int world,
                                           int64_t dst_ld, int row0, void* stream) {
  DRB_REQUIRE(qkv && wq && wk && cos_tab && sin_tab && dst_ptrs, "null pointer");
  DRB_REQUIRE(S_local > 0 && num_heads > 0 && row0 >= 0, "bad sizes");
  DRB_REQUIRE(world >= 1 && world <= DRB_CP_MAX_RANKS && num_heads % world == 0, "world must divide num_heads (<= 8 ranks)");
  DRB_REQUIRE(ld % 4 == 0 && ld >= 3LL * num_heads * 128, "qkv pitch must hold [q | k | v] with 128-wide heads");
  DRB_REQUIRE(dst_ld % 4 == 0 && dst_ld >= 3LL * (num_heads / world) * 128, "destination pitch too small");
  PeerPtrs d{};
  for (int i = 0; i < world; ++i) {
    DRB_REQUIRE(dst_ptrs[i] != nullptr && (reinterpret_cast<uintptr_t>(dst_ptrs[i]) & 7) == 0, "bad destination pointer");
    d.p[i] = dst_ptrs[i];
  }
  qk_norm_rope_scatter_kernel<<<S_local, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), ld, static_cast<const __nv_bfloat16*>(wq), static_cast<const __nv_bfloat16*>(wk),
      static_cast<const __nv_bfloat16*>(cos_tab), static_cast<const __nv_bfloat16*>(sin_tab), num_heads, d, dst_ld, row0,
      num_heads / world);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int drb_cp_barrier(void* const* flag_ptrs, int rank, int world, uint32_t epoch, void* stream) {
  DRB_REQUIRE(flag_ptrs != nullptr, "null pointer");
  DRB_REQUIRE(world >= 1 && world <= DRB_CP_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world");
  PeerPtrs f{};
  for (int i = 0; i < world; ++i) {
    DRB_REQUIRE(flag_ptrs[i] != nullptr, "null flag pointer");
    f.p[i] = flag_ptrs[i];
  }
  cp_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, rank, world, epoch);
  DRB_CUDA(cudaGetLastError());
  return 0;
}

// ---- peer memory: cudaMalloc + CUDA IPC (one process per GPU) --------------------------------------------------
extern "C" int drb_peer_alloc(int64_t bytes, void** ptr) {
  DRB_REQUIRE(ptr != nullptr && bytes > 0, "bad arguments");
  DRB_CUDA(cudaMalloc(ptr, static_cast<size_t>(bytes)));
  DRB_CUDA(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)));
  return 0;
}

extern "C" int drb_peer_free(void* ptr) {
  DRB_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int drb_peer_export(const void* ptr, void* handle64) {
  DRB_REQUIRE(ptr && handle64, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == DRB_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  DRB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

extern "C" int drb_peer_import(const void* handle64, void** ptr) {
  DRB_REQUIRE(ptr && handle64, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  DRB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int drb_peer_close(void* ptr) {
  DRB_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}
