// Micro-benchmark: issue throughput of MUFU.EX2, packed fma.rn.f32x2, FMNMX, cvt.bf16x2 per SM sub-partition on sm_100a.
// One block per SM, W warps per block (W/4 per SMSP); each thread runs N independent chains.  Prints cycles per
// warp-instruction per SMSP.   nvcc -arch=sm_100a -O3 -o sfu sfu.cu && ./sfu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, unsigned long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i * 0.1f - 3.0f;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
    }
    if (OP == 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        unsigned long long v;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(v));
      }
    }
    if (OP == 4) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        unsigned int r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i + 1]));
        a[i] = __uint_as_float(r);
      }
    }
  }
  unsigned long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMallocManaged(&cyc, 8);
  const char* names[] = {"ex2.approx.f32", "fma.rn.f32", "fma.rn.f32x2 (x2 per pair, counted per packed instr)", "max.f32", "cvt.rn.bf16x2.f32"};
  const int per_iter[] = {8, 8, 8, 8, 4};
  const int iters = 4096;
  for (int op = 0; op < 5; ++op) {
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
        if (op == 4) k<4><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      double instr_per_smsp = (double)iters * per_iter[op] * (warps / 4);
      printf("%-55s warps/SMSP=%d : %.2f cycles per warp-instruction per SMSP\n", names[op], warps / 4, *cyc / instr_per_smsp);
    }
  }
  return 0;
}
