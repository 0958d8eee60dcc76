// Throughput / latency of the attention kernel's softmax inner step in isolation: cycles per 32-element chunk per warp for
// several FMA-pipe fractions and 1/2/4 warps per SM sub-partition.  Inputs come from shared memory and P goes back to
// shared memory each iteration (standing in for tcgen05.ld / tcgen05.st), so nothing is loop-invariant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o softmax_chunk softmax_chunk.cu ../../diffusionrenderer-comfyui_b200/csrc/runtime.cu
#include <cstdio>
#include "../../diffusionrenderer-comfyui_b200/csrc/attention.cu"

using namespace drb;

template <uint32_t kMask>
__global__ void bench(float* out, int iters, unsigned long long* cyc) {
  extern __shared__ uint4 sm[];                     // [blockDim.x][8] uint4 = 32 floats per thread
  uint4* mine = sm + threadIdx.x;                   // element k at mine[k * blockDim.x] (conflict-free)
  for (int k = 0; k < 8; ++k)
    mine[k * blockDim.x] = make_uint4(__float_as_uint(-0.01f * ((threadIdx.x + k) % 97)), __float_as_uint(-0.3f), __float_as_uint(-1.1f * k),
                                      __float_as_uint(-2.0f));
  __syncthreads();
  const uint64_t scale2 = pack2(kScaleLog2, kScaleLog2), negm2 = pack2(-0.5f, -0.5f);
  uint64_t sum = pack2(0.f, 0.f);
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t s[32], pk[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint4 v = mine[k * blockDim.x];
      s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w;
    }
    softmax_chunk<kMask>(s, scale2, negm2, pk, sum);
#pragma unroll
    for (int k = 0; k < 4; ++k)                      // feed P back so that the next iteration's inputs depend on it
      mine[k * blockDim.x] = make_uint4(pk[4 * k] | 0x80008000u, pk[4 * k + 1] | 0x80008000u, pk[4 * k + 2] | 0x80008000u, pk[4 * k + 3] | 0x80008000u);
  }
  unsigned long long t1 = clock64();
  float a, b;
  unpack2(sum, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <uint32_t kMask>
void run(const char* name, float* out, unsigned long long* cyc) {
  const int iters = 2000;
  cudaFuncSetAttribute(bench<kMask>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 128);
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      bench<kMask><<<148, warps * 32, warps * 32 * 128>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    printf("%-12s warps/SMSP=%d : %7.1f cycles per chunk per warp, %7.1f per chunk per SMSP\n", name, warps / 4,
           (double)*cyc / iters, (double)*cyc / iters / (warps / 4));
  }
}

int main() {
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMallocManaged(&cyc, 8);
  run<0x0000u>("poly 0/16", out, cyc);
  run<0x0101u>("poly 2/16", out, cyc);
  run<0x0421u>("poly 3/16", out, cyc);
  run<0x1111u>("poly 4/16", out, cyc);
  run<0x4924u>("poly 5/16", out, cyc);
  run<0x5555u>("poly 8/16", out, cyc);
  return 0;
}
