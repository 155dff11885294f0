// Micro-benchmark: scalar FFMA/FADD versus packed FFMA2/FADD2 (sm_100a f32x2) issue + pipe throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, float s) {
  float a[16]; unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const unsigned long long ss = pk(s, s);
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, ss);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], ss);
    } else if (MODE == 4) {   // 8 FADD2 + 8 integer ops: does packing free issue slots for other pipes?
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = add2(p[i], ss); asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+f"(a[i]) : "f"(s)); }
    } else if (MODE == 5) {   // 16 FADD + 8 integer ops
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(*(unsigned*)&p[i]) : "r"(__float_as_uint(s)));
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)(p[i] & 0xffffffffu)) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, double flop_per_thread_iter, float* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8;
  k<MODE><<<grid, 256>>>(d, 1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = flop_per_thread_iter * ITERS * 256.0 * grid;
  printf("%-28s %8.3f ms  %8.2f TFLOP/s (or Tops/s)\n", name, ms, fl / ms / 1e9);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("FFMA  x16", 32, d); run<1>("FFMA2 x8", 32, d); run<2>("FADD  x16", 16, d); run<3>("FADD2 x8", 16, d);
  run<4>("FADD2 x8 + LOP3 x8", 16, d); run<5>("FADD x16 + LOP3 x8", 16, d);
  return 0;
}
