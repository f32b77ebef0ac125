// Micro-benchmark: MUFU.EX2 throughput per SM for f32 / bf16x2 / f16x2 forms (hard part 1).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mufu_bench tools/mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; u[i] = __float_as_uint(a[i]) | 0x3c003c00u; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int elems_per_op) {
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 4096, blocks = 148 * 8, threads = 256;
  k<MODE><<<blocks, threads>>>(out, 16, 0.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ops = (double)blocks * threads * iters * 8;
  double per_s = ops / (ms * 1e-3);
  printf("%-22s %8.3f ms  %.2f Tinstr-lanes/s  -> %.2f lane-ops/clk/SM @%d MHz (x%d elems/op = %.2f elems/clk/SM)\n", name, ms,
         per_s / 1e12, per_s / 148 / (clk_khz * 1e3), clk_khz / 1000, elems_per_op, elems_per_op * per_s / 148 / (clk_khz * 1e3));
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("fma.rn.f32", 1);
  return 0;
}
