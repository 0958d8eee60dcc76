// Cross-GPU ordering folded INTO the kernels of the context-parallel exchange: instead of a stand-alone
// barrier kernel between the producer and the consumer of peer-stored data (which drains the GPU twice per transformer
// block), the producer kernel's last CTA publishes a flag to every peer once all of its P2P stores are fenced, and the
// consumer kernel's TMA-producer thread waits for the flags of all ranks right before its first operand load — its
// prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps the tail of the peers' producers.
//
// Flag memory of rank j (peer-mapped, zeroed): uint32 [DRB_CP_FLAG_SLOTS][DRB_CP_MAX_RANKS]; word [slot][r] is written by
// rank r only.  Epochs increase monotonically per slot and are the same on every rank (SPMD launch order).  Word
// [DRB_CP_STATUS_WORD] of the LOCAL array is set to 1 if a wait timed out (the host checks it after synchronising; the
// kernel then proceeds on whatever data is there instead of killing the CUDA context with a trap).
#pragma once
#include <stdint.h>

#include "../../include/drb200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace drb {

struct CpSync {
  void* flag_ptrs[DRB_CP_MAX_RANKS];
  uint32_t* counter;      // local, zero between uses: counts the CTAs of the producer kernel that finished their stores
  int world, rank;
  int signal_slot, wait_slot;
  uint32_t signal_epoch, wait_epoch;   // 0 = no signal / no wait
  uint32_t timeout_ms;
};

// one thread; returns after flags [wait_slot][r] >= wait_epoch for every rank r (or after the timeout)
__device__ __forceinline__ void cp_wait(const CpSync& s) {
  if (s.wait_epoch == 0) return;
  const uint32_t* mine = static_cast<const uint32_t*>(s.flag_ptrs[s.rank]) + s.wait_slot * DRB_CP_MAX_RANKS;
  const uint64_t t0 = globaltimer_ns();
  for (int r = 0; r < s.world; ++r) {
    uint32_t v, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + r) : "memory");
      if (static_cast<int32_t>(v - s.wait_epoch) >= 0) break;
      if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > static_cast<uint64_t>(s.timeout_ms) * 1000000ull) {
        static_cast<uint32_t*>(s.flag_ptrs[s.rank])[DRB_CP_STATUS_WORD] = 1u;   // a lost peer: report, do not hang
        return;
      }
    }
  }
  // the operands were written by other GPUs' generic-proxy stores; they are read next by TMA (async proxy)
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// one thread per CTA, called after a CTA-level barrier that follows the storing threads' __threadfence_system():
// the last CTA of the grid publishes [signal_slot][rank] = signal_epoch on every rank
__device__ __forceinline__ void cp_signal_when_grid_done(const CpSync& s, uint32_t total_ctas) {
  if (s.signal_epoch == 0) return;
  __threadfence_system();
  const uint32_t prev = atomicAdd(s.counter, 1u);
  if (prev + 1 == total_ctas) {
    atomicExch(s.counter, 0u);
    __threadfence_system();
    for (int j = 0; j < s.world; ++j) {
      uint32_t* remote = static_cast<uint32_t*>(s.flag_ptrs[j]) + s.signal_slot * DRB_CP_MAX_RANKS + s.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(s.signal_epoch) : "memory");
    }
  }
}

// host: validate the C-ABI descriptor and copy it into the kernel-parameter form (a null descriptor = no sync)
inline int fill_cp_sync(CpSync* out, const drb_cp_sync* in) {
  *out = CpSync{};
  if (in == nullptr) return 0;
  DRB_REQUIRE(in->flag_ptrs != nullptr && in->world >= 1 && in->world <= DRB_CP_MAX_RANKS && in->rank >= 0 && in->rank < in->world,
              "bad rank / world in the sync descriptor");
  DRB_REQUIRE(in->signal_slot >= 0 && in->signal_slot < DRB_CP_FLAG_SLOTS && in->wait_slot >= 0 && in->wait_slot < DRB_CP_FLAG_SLOTS,
              "flag slot out of range");
  DRB_REQUIRE(in->signal_epoch == 0 || (in->counter != nullptr && (reinterpret_cast<uintptr_t>(in->counter) & 3) == 0),
              "a signalling kernel needs the CTA counter");
  for (int i = 0; i < in->world; ++i) {
    DRB_REQUIRE(in->flag_ptrs[i] != nullptr && (reinterpret_cast<uintptr_t>(in->flag_ptrs[i]) & 3) == 0, "bad flag pointer");
    out->flag_ptrs[i] = in->flag_ptrs[i];
  }
  out->counter = in->counter;
  out->world = in->world;
  out->rank = in->rank;
  out->signal_slot = in->signal_slot;
  out->wait_slot = in->wait_slot;
  out->signal_epoch = in->signal_epoch;
  out->wait_epoch = in->wait_epoch;
  out->timeout_ms = in->timeout_ms ? in->timeout_ms : 60000u;
  return 0;
}

}  // namespace drb
