// Micro-benchmark: issue throughput per SM of the instructions in the softmax inner loops, alone and mixed, to see which
// share a pipe with MUFU.EX2 (cvt.rn.bf16x2.f32 = F2FP, fmul, fmax3, integer rounding + prmt packing).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pipe_bench tools/pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 5 || MODE == 9) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1 || MODE == 2) {
        unsigned t;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[i]), "f"(__uint_as_float(u[i])));
        u[i] ^= t;                                   // keep the result live (one LOP3 per cvt)
      }
      if (MODE == 10) u[i] ^= __float_as_uint(a[i]) + u[(i + 1) & 7];   // LOP3 + IADD baseline
      if (MODE == 3 || MODE == 4) asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (MODE == 5) {   // ex2 + 2 fma (polynomial-ish filler)
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 3) & 7]) : "f"(seed));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 5) & 7]) : "f"(seed));
      }
      if (MODE == 6) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (MODE == 7) {
        unsigned t;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[i]), "f"(__uint_as_float(u[i])));
        u[i] ^= t;
      }
      if (MODE == 8 || MODE == 9) {   // integer round-half-up to bf16 + pack: 2 iadd + 1 prmt per pair
        unsigned x = __float_as_uint(a[i]) + 0x8000u, y = __float_as_uint(a[(i + 1) & 7]) + 0x8000u;
        unsigned t;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(t) : "r"(x), "r"(y ^ u[i]));
        u[i] ^= t;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int ops_per_iter) {
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
  const double groups = (double)blocks * threads * iters * 8;   // one "group" = the instructions of one unrolled slot
  printf("%-34s %8.3f ms  %7.2f groups/clk/SM  (%d instr per group -> %.2f instr-lanes/clk/SM)\n", name, ms,
         groups / (ms * 1e-3) / 148 / (clk_khz * 1e3), ops_per_iter, ops_per_iter * groups / (ms * 1e-3) / 148 / (clk_khz * 1e3));
}

int main() {
  run<0>("ex2", 1);
  run<1>("cvt.rn.bf16x2.f32", 1);
  run<7>("cvt.rn.f16x2.f32", 1);
  run<2>("ex2 + cvt.bf16x2", 2);
  run<4>("fmul", 1);
  run<3>("ex2 + fmul", 2);
  run<5>("ex2 + 2 fma", 3);
  run<6>("max3", 1);
  run<8>("2 iadd + prmt (int pack)", 3);
  run<9>("ex2 + 2 iadd + prmt", 4);
  run<10>("lop3 + iadd baseline", 2);
  return 0;
}
