// Host-side helpers shared by the .cu files: error plumbing for the C ABI and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>

namespace drb {

// ---- error state of the C ABI (thread-local; read back with drb_last_error()) ----
void set_error(const std::string& msg);
int fail(const char* where, const std::string& msg);             // sets error, returns DRB_ERR_*
int check_cuda(cudaError_t e, const char* where);                // 0 on success

#define DRB_CUDA(expr)                                        \
  do {                                                        \
    int _rc = ::drb::check_cuda((expr), #expr);               \
    if (_rc != 0) return _rc;                                 \
  } while (0)

#define DRB_REQUIRE(cond, msg)                                \
  do {                                                        \
    if (!(cond)) return ::drb::fail(__func__, (msg));         \
  } while (0)

int num_sms();   // multiprocessor count of the current device (cached)

// One-time kernel configuration PER DEVICE: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) belongs to the device's
// context, so a process that drives several GPUs (ComfyUI with more than one device, torch.cuda.set_device switching)
// must configure every kernel on each of them.  `configure` runs once per device index under a mutex; a launch on
// another thread waits for it instead of racing ahead with the 48 KB default.
struct DeviceOnce {
  std::atomic<uint64_t> done{0};
  std::mutex mu;
};
template <typename F>
int device_once(DeviceOnce& st, F&& configure) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  if (dev < 0 || dev >= 64) return configure();   // unknown index: configure on every call (cheap, idempotent)
  const uint64_t bit = 1ull << dev;
  if (st.done.load(std::memory_order_acquire) & bit) return 0;
  std::lock_guard<std::mutex> lock(st.mu);
  if (st.done.load(std::memory_order_relaxed) & bit) return 0;
  const int rc = configure();
  if (rc == 0) st.done.fetch_or(bit, std::memory_order_release);
  return rc;
}

// ---- TMA descriptors -------------------------------------------------------------------
// 2-D bf16 row-major tensor [rows][cols] with a row pitch of `ld` elements; box = box_rows x box_cols,
// 128-byte swizzle (box_cols * 2 must be 128).  Returns 0 or an error code.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);
// rank-`rank` (2..5) bf16 tensor, innermost dimension first: dims[rank] in elements, strides[rank-1] in bytes (dimension 0
// is contiguous), box[rank] in elements with box[0]*2 == 128; 128-byte swizzle, out-of-bounds elements read as zero.
int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box);

}  // namespace drb
