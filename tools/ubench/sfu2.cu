// more issue-rate probes: max.bf16x2, 3-input max.f32, add.rn.f32x2, ex2.approx.f16x2/bf16x2, lop3/shift
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, unsigned long long* cyc) {
  unsigned int a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3f803f80u + threadIdx.x + i;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
      if (OP == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&a[(i + 1) & 7]), "f"(*(float*)&a[(i + 2) & 7]));
      if (OP == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (OP == 3) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (OP == 4) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&a[(i + 1) & 7]));
      if (OP == 5) asm volatile("shl.b32 %0, %0, 1;" : "+r"(a[i]));
      if (OP == 6) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(*(float*)&a[i]), "f"(*(float*)&a[(i + 1) & 7]));
      if (OP == 7) asm volatile("fma.rn.bf16x2 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
    }
  }
  unsigned long long t1 = clock64();
  unsigned int s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMallocManaged(&cyc, 8);
  const char* names[] = {"max.bf16x2", "max.f32 (3-input)", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "add.rn.f32", "shl.b32", "cvt.rn.f16x2.f32", "fma.rn.bf16x2"};
  const int iters = 4096;
  for (int op = 0; op < 8; ++op)
    for (int warps = 4; warps <= 16; warps *= 4) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: k<0><<<148, warps * 32>>>(out, iters, cyc); break;
          case 1: k<1><<<148, warps * 32>>>(out, iters, cyc); break;
          case 2: k<2><<<148, warps * 32>>>(out, iters, cyc); break;
          case 3: k<3><<<148, warps * 32>>>(out, iters, cyc); break;
          case 4: k<4><<<148, warps * 32>>>(out, iters, cyc); break;
          case 5: k<5><<<148, warps * 32>>>(out, iters, cyc); break;
          case 6: k<6><<<148, warps * 32>>>(out, iters, cyc); break;
          case 7: k<7><<<148, warps * 32>>>(out, iters, cyc); break;
        }
        cudaDeviceSynchronize();
      }
      printf("%-28s warps/SMSP=%d : %.2f cycles per warp-instruction per SMSP\n", names[op], warps / 4, *cyc / ((double)iters * 8 * (warps / 4)));
    }
  return 0;
}
