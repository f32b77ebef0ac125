// Micro-benchmark: TMEM read/write throughput per SM (tcgen05.ld / tcgen05.st) and how it overlaps MUFU.EX2.
// Sizes the softmax loops of the attention kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_X32(taddr, r)                                                                                                  \
  asm volatile(                                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                           \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"  \
      "%30,%31}, [%32];"                                                                                                  \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),       \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),            \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),           \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                         \
      : "r"(taddr)                                                                                                        \
      : "memory")
#define ST_X32(taddr, r)                                                                                                  \
  asm volatile(                                                                                                           \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                     \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"  \
      "%31,%32};" ::"r"(taddr),                                                                                           \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),       \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),         \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),         \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                      \
      : "memory")

// MODE 0: ld only (wait after every `depth` loads); 1: st only; 2: ld + 32 ex2 per load; 3: 32 ex2 only (same loop);
// 4: ld + 32 ex2 + 16 cvt.bf16x2 + st x16-equivalent (x32 st of half the regs every other iteration)
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cycles) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;   // warps 4-7 use columns 128..
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(0.001f * (threadIdx.x + i));
  ST_X32(tl, r);
  ST_X32(tl + 32, r);
  ST_X32(tl + 64, r);
  ST_X32(tl + 96, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (MODE == 0 || MODE == 2 || MODE == 4) LD_X32(tl + ((it * DEPTH + d) & 3) * 32, r);
      if (MODE == 0 && d == DEPTH - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (MODE == 2 || MODE == 4) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (MODE == 2 || MODE == 3 || MODE == 4) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float f = __uint_as_float(r[i]);
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
          if (MODE == 3) r[i] = __float_as_uint(f * 0.5f); else acc += f;
        }
      }
      if (MODE == 1) ST_X32(tl + ((it * DEPTH + d) & 3) * 32, r);
      if (MODE == 1 && d == DEPTH - 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (MODE == 0) acc += __uint_as_float(r[it & 31]);
  }
  const long long t1 = clock64();
  if (MODE == 3) for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
  if (threadIdx.x == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) out[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

template <int MODE, int DEPTH>
void run(const char* name, int threads, int ctas_per_sm) {
  float* out;
  long long* cyc;
  const int blocks = 148 * ctas_per_sm, iters = 2048;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  k<MODE, DEPTH><<<blocks, threads>>>(out, 16, nullptr);
  k<MODE, DEPTH><<<blocks, threads>>>(out, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148 * 4];
  cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  const double ops = (double)iters * DEPTH;                     // x32 ops per warp
  const double bytes_per_sm = ops * (threads / 32) * ctas_per_sm * 4096.0;
  printf("%-44s thr=%3d cta/sm=%d  %8.1f cyc/op/warp  -> %7.1f B/clk/SM (TMEM bytes)  %6.2f elems/clk/SM\n", name, threads,
         ctas_per_sm, avg / ops, bytes_per_sm / avg, bytes_per_sm / 4 / avg);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0, 1>("ld x32, wait each", 128, 1);
  run<0, 4>("ld x32, wait every 4", 128, 1);
  run<0, 4>("ld x32, wait every 4", 256, 1);
  run<0, 4>("ld x32, wait every 4", 128, 2);
  run<1, 4>("st x32, wait every 4", 128, 1);
  run<1, 4>("st x32, wait every 4", 256, 1);
  run<3, 1>("32 ex2 only", 128, 1);
  run<3, 1>("32 ex2 only", 256, 1);
  run<2, 1>("ld x32 + wait + 32 ex2", 128, 1);
  run<2, 1>("ld x32 + wait + 32 ex2", 256, 1);
  run<2, 1>("ld x32 + wait + 32 ex2", 128, 2);
  return 0;
}
