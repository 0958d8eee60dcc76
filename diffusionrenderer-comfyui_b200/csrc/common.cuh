// Host-side helpers shared by the .cu files: error plumbing for the C ABI and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

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

// ---- TMA descriptors -------------------------------------------------------------------
// 2-D bf16 row-major tensor [rows][cols] with a row pitch of `ld` elements; box = box_rows x box_cols,
// 128-byte swizzle (box_cols * 2 must be 128).  Returns 0 or an error code.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);
// 4-D bf16 channels-last activation [T][H][W][C] (C contiguous); box = (bt, bh, bw, bc) with bc*2 == 128.
int make_tmap_4d_bf16(CUtensorMap* out, const void* base, uint64_t T, uint64_t H, uint64_t W, uint64_t C,
                      uint32_t bt, uint32_t bh, uint32_t bw, uint32_t bc);

}  // namespace drb
