// Host-side runtime of libdrb200.so: thread-local error state, device queries and TMA tensor-map encoding.
// No kernels live here.
#include <cudaTypedefs.h>

#include <mutex>

#include "../../include/drb200.h"
#include "common.cuh"

namespace drb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(const char* where, const std::string& msg) {
  g_last_error = std::string(where) + ": " + msg;
  return DRB_ERR_INVALID;
}

int check_cuda(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  g_last_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return DRB_ERR_CUDA;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so the library has no
// link-time dependency on libcuda (it must load on a box without a driver for the symbol-export test).
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  auto fn = encode_fn();
  if (!fn) {
    g_last_error = "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)";
    return DRB_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail("make_tmap_2d_bf16", "base pointer not 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail("make_tmap_2d_bf16", "row pitch must be a multiple of 8 bf16 elements");
  if (box_cols * 2 != 128 || box_rows == 0 || box_rows > 256) return fail("make_tmap_2d_bf16", "bad box");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_error = "cuTensorMapEncodeTiled(2d) failed with CUresult " + std::to_string(static_cast<int>(r));
    return DRB_ERR_CUDA;
  }
  return 0;
}

int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box) {
  auto fn = encode_fn();
  if (!fn) {
    g_last_error = "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)";
    return DRB_ERR_CUDA;
  }
  if (rank < 2 || rank > 5) return fail("make_tmap_nd_bf16", "rank must be 2..5");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail("make_tmap_nd_bf16", "base pointer not 16-byte aligned");
  if (box[0] * 2 != 128) return fail("make_tmap_nd_bf16", "box must cover 64 innermost elements (128 bytes)");
  cuuint64_t d[5];
  cuuint64_t st[4];
  cuuint32_t b[5];
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    if (dims[i] == 0 || box[i] == 0 || box[i] > 256) return fail("make_tmap_nd_bf16", "bad dimension or box");
    d[i] = dims[i];
    b[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides[i] % 16 != 0) return fail("make_tmap_nd_bf16", "strides must be multiples of 16 bytes");
    st[i] = strides[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, st, b, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_error = "cuTensorMapEncodeTiled(nd) failed with CUresult " + std::to_string(static_cast<int>(r));
    return DRB_ERR_CUDA;
  }
  return 0;
}

}  // namespace drb

extern "C" {

const char* drb_last_error(void) { return drb::g_last_error.c_str(); }

int drb_version(void) { return 100; }

int drb_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

}  // extern "C"
