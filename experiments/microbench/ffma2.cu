// Packed fp32 (FFMA2 / FADD2, PTX fma.rn.f32x2) against scalar FFMA on sm_100a: issue throughput per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2.cu -o ffma2 && ./ffma2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void k(float* out, float s, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  uint64_t p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = ((uint64_t)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
  const uint64_t s2 = ((uint64_t)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], s2, s2);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 2, 1024>>>(out, 0.999f, iters); else k<1><<<148 * 2, 1024>>>(out, 0.999f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)148 * 2 * 1024 * iters * 16;   // scalar-equivalent FMAs
      printf("%s: %.3f ms, %.1f TFMA/s (scalar-equivalent), %.1f FMA/clk/SM at 1.9 GHz\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9,
             fma / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
