// Throughput of the attention kernel's softmax inner step in isolation: cycles per 32-element chunk per warp, for several
// fractions of FMA-pipe exponentials and 1/2/4 warps per SM sub-partition.  Uses the kernel's own device functions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o softmax_chunk softmax_chunk.cu ../../diffusionrenderer-comfyui_b200/csrc/runtime.cu
#include <cstdio>
#include "../../diffusionrenderer-comfyui_b200/csrc/attention.cu"

using namespace drb;

template <uint32_t kMask>
__global__ void bench(float* out, int iters, unsigned long long* cyc) {
  uint32_t s[32];
  for (int i = 0; i < 32; ++i) s[i] = __float_as_uint(-0.01f * ((threadIdx.x * 7 + i * 13) % 97));
  const uint64_t scale2 = pack2(kScaleLog2, kScaleLog2), negm2 = pack2(-0.5f, -0.5f);
  uint64_t sum = pack2(0.f, 0.f);
  float mx = -INFINITY;
  uint32_t acc = 0;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
    softmax_chunk<kMask>(s, scale2, negm2, pk, sum);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] ^= ((acc >> (i & 15)) & 1u);   // every input changes every iteration (32 LOP3 of overhead)
  }
  unsigned long long t1 = clock64();
  float a, b;
  unpack2(sum, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + mx + __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <uint32_t kMask>
void run(const char* name, float* out, unsigned long long* cyc) {
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      bench<kMask><<<148, warps * 32>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    printf("%-22s warps/SMSP=%d : %7.1f cycles per chunk per warp, %7.1f per chunk per SMSP\n", name, warps / 4,
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
