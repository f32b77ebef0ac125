// Micro-benchmark: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, cta_group::1, M=128) cycles per instruction as a function of
// N, operand source (SS / TS), B layout (K-major / MN-major) and accumulator dependence (same accumulator back to back
// vs round-robin over several).  Sizes the MMA issue order of the attention kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I chexpert_b200/csrc -o tools/bin/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace aaconv;

struct __align__(1024) Smem {
  __nv_bfloat16 a[128 * 64];      // 16 KB: one K-major atom, 128 rows
  __nv_bfloat16 b[4][256 * 64];   // up to N=256 rows (K-major) / 4 atoms of 64 columns (MN-major)
  uint64_t bar;
  uint32_t tmem_base;
};

// mode: 0 SS K-major B; 1 TS (A from TMEM) K-major B; 2 TS MN-major B; 3 SS MN-major B
__global__ void __launch_bounds__(288, 1) k(int mode, int N, int nacc, int iters, long long* cycles, int hammer) {
  extern __shared__ __align__(1024) uint8_t raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (int)(sizeof(Smem) / 4) - 8; i += blockDim.x) reinterpret_cast<uint32_t*>(&sm)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) { tc::mbar_init(&sm.bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (warp == 0) {                                       // whole warp runs the uniform loop, one elected lane issues
    const bool bmn = mode >= 2;
    const uint32_t idesc = tc::idesc_bf16_f32(128, N) | (bmn ? (1u << 16) : 0u);
    const uint32_t a_lo = tc::desc_lo_k(tc::smem_u32(sm.a));
    const uint32_t b_lo = bmn ? tc::desc_lo_mn(tc::smem_u32(sm.b[0]), 64 * 128) : tc::desc_lo_k(tc::smem_u32(sm.b[0]));
    const uint32_t accw = (N + 31) & ~31;                // accumulator stride in columns
    const uint32_t a_tmem = tmem + 448;                  // TS: A operand columns (8 per k-step), outside the accumulators
    const uint32_t bstep = bmn ? 128 : 2;
    const uint32_t d1 = tmem + (nacc > 1 ? accw : 0), d2 = tmem + (nacc > 2 ? 2 * accw : 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it += 12) {            // 12 MMAs per trip: k-steps 0..3, accumulators round robin
      if (tc::elect_one()) {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const uint32_t d = (u % 3 == 0) ? tmem : (u % 3 == 1 ? d1 : d2);
          if (mode == 0 || mode == 3) tc::mma_ss(d, tc::desc64(a_lo + (u & 3) * 2), tc::desc64(b_lo + (u & 3) * bstep), idesc, 1);
          else tc::mma_ts(d, a_tmem + (u & 3) * 8, tc::desc64(b_lo + (u & 3) * bstep), idesc, 1);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (tc::elect_one()) tc::mma_commit(&sm.bar);
    __syncwarp();
    tc::mbar_wait(&sm.bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      cycles[blockIdx.x * 2] = t1 - t0;
      cycles[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  else if (hammer) {
    // 8 more warps (two per TMEM lane quadrant) read TMEM concurrently: hammer 1 = ld only, 2 = ld + 32 ex2 per load
    __shared__ volatile int stop;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp >> 2) & 1) * 64;
    uint32_t r[32];
    float acc = 0.f;
    for (int it = 0; it < iters * 4; ++it) {
      tc::tmem_ld_x32(tl + (it & 1) * 32, r);
      tc::tmem_ld_wait();
      if (hammer == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += tc::ex2f(__uint_as_float(r[i]));
      } else acc += __uint_as_float(r[it & 31]);
    }
    if (acc == 123.456f) cycles[0] = 0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// The issue pattern of the dq kernel's MMA warp: per tile 8 score MMAs (TS, N=64) + commit, then 4 gradient MMAs
// (TS, MN-major B, N=112) + commit; flags switch the fences / commits off to see what they cost.
template <int VAR, int NS_, int NG_>
__global__ void __launch_bounds__(384, 1) kpat2(int tiles, long long* cycles, int hammer) {
  extern __shared__ __align__(1024) uint8_t raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t sink;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (int)(sizeof(Smem) / 4) - 8; i += blockDim.x) reinterpret_cast<uint32_t*>(&sm)[i] = 0;
  __syncthreads();
  __shared__ uint64_t done_bar;
  if (threadIdx.x == 0) { tc::mbar_init(&sm.bar, 1); tc::mbar_init(&sink, 1 << 20); tc::mbar_init(&done_bar, 1); tc::fence_barrier_init(); tc::mbar_arrive(&done_bar); }
  if (warp == 0) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (warp == 0 && (VAR != 2 || lane == 0)) {
    constexpr uint32_t idesc_s = tc::idesc_bf16_f32(128, 64), idesc_g = tc::idesc_bf16_f32(128, 112) | (1u << 16);
    const uint32_t bk = tc::desc_lo_k(tc::smem_u32(sm.b[0])), bm = tc::desc_lo_mn(tc::smem_u32(sm.b[1]), 64 * 128);
    const bool leader = VAR == 2 ? true : tc::elect_one();
    const long long t0 = clock64();
    for (int j = 0; j < tiles; ++j) {
      const uint32_t slot = tmem + 64 + 128 * (j & 1);
      if (VAR >= 3) tc::mbar_wait(&done_bar, 0);
      if (VAR != 4) tc::tc_fence_after();
      if (VAR == 0 ? tc::elect_one() : leader) {
#pragma unroll
        for (int ks = 0; ks < NS_; ++ks) tc::mma_ts(slot + (ks == NS_ - 1 ? 64 : 0), tmem + (ks & 7) * 8, tc::desc64(bk + (ks & 3) * 2), idesc_s, (ks > 0 && ks < NS_ - 1) ? 1u : 0u);
        tc::mma_commit(&sink);
      }
      if (VAR == 0) __syncwarp();
      if (VAR >= 3) tc::mbar_wait(&done_bar, 0);
      if (VAR != 4) tc::tc_fence_after();
      if (VAR == 0 ? tc::elect_one() : leader) {
#pragma unroll
        for (int ks = 0; ks < NG_; ++ks) tc::mma_ts(tmem + 320, slot + 64 + ks * 8, tc::desc64(bm + ks * 128), idesc_g, 1);
        tc::mma_commit(&sink);
        tc::mma_commit(&sink);
      }
      if (VAR == 0) __syncwarp();
    }
    const long long t1 = clock64();
    if (leader) tc::mma_commit(&sm.bar);
    tc::mbar_wait(&sm.bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { cycles[blockIdx.x * 2] = t1 - t0; cycles[blockIdx.x * 2 + 1] = t2 - t0; }
  } else if (warp >= 4 && hammer) {
    // 8 "math" warps imitating the dq kernel's per-tile work on TMEM columns the MMAs do not touch:
    // hammer bit0: 4 x tcgen05.ld.x32, bit1: 64 ex2 + 64 fmul + 32 cvt, bit2: 2 x tcgen05.st.x16 per tile
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 448;
    uint32_t r[32], q[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(0.001f * (lane + i));
    float acc = 0.f;
    for (int j = 0; j < tiles * 2; ++j) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (hammer & 1) { tc::tmem_ld_x32(tl, r); tc::tmem_ld_x32(tl + 32, r); tc::tmem_ld_wait(); }
        if (hammer & 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = tc::ex2f(__uint_as_float(r[2 * i]) * 0.001f), p1 = tc::ex2f(__uint_as_float(r[2 * i + 1]) * 0.001f);
            q[i] = tc::pack_bf16x2(p0 * acc, p1 * acc);
            acc += p0;
          }
        }
        if (hammer & 4) tc::tmem_st_x16(tl + 32 + c * 16, q);
      }
      if (hammer & 4) tc::tmem_st_wait();
    }
    if (acc == 123.456f) cycles[0] = q[3];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

template <int VAR, int NS_, int NG_>
static void runpat2(const char* name, int hammer = 0) {
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * 2 * 148);
  const size_t smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(kpat2<VAR, NS_, NG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 50;
  kpat2<VAR, NS_, NG_><<<148, hammer ? 384 : 128, smem>>>(4, cyc, hammer);
  kpat2<VAR, NS_, NG_><<<148, hammer ? 384 : 128, smem>>>(tiles, cyc, hammer);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[2 * 148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < 148; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  printf("%-52s issue %7.1f  complete %7.1f cyc/tile\n", name, issue / 148 / tiles, total / 148 / tiles);
  cudaFree(cyc);
}

__global__ void __launch_bounds__(128, 1) kpat(int tiles, int use_fence, int use_commit, int nS, int nG, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t sink;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (int)(sizeof(Smem) / 4) - 8; i += blockDim.x) reinterpret_cast<uint32_t*>(&sm)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) { tc::mbar_init(&sm.bar, 1); tc::mbar_init(&sink, 1 << 20); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (warp == 0) {
    const uint32_t idesc_s = tc::idesc_bf16_f32(128, 64), idesc_g = tc::idesc_bf16_f32(128, 112) | (1u << 16);
    const uint32_t bk = tc::desc_lo_k(tc::smem_u32(sm.b[0])), bm = tc::desc_lo_mn(tc::smem_u32(sm.b[1]), 64 * 128);
    const long long t0 = clock64();
    for (int j = 0; j < tiles; ++j) {
      const uint32_t slot = tmem + 64 + 128 * (j & 1);
      if (use_fence) tc::tc_fence_after();
      if (tc::elect_one()) {
        for (int ks = 0; ks < nS; ++ks) tc::mma_ts(slot + (ks == nS - 1 ? 64 : 0), tmem + (ks & 7) * 8, tc::desc64(bk + (ks & 3) * 2), idesc_s, (ks > 0 && ks < nS - 1) ? 1u : 0u);
        if (use_commit) tc::mma_commit(&sink);
      }
      __syncwarp();
      if (use_fence) tc::tc_fence_after();
      if (tc::elect_one()) {
        for (int ks = 0; ks < nG; ++ks) tc::mma_ts(tmem + 320, slot + 64 + ks * 8, tc::desc64(bm + ks * 128), idesc_g, 1);
        if (use_commit) { tc::mma_commit(&sink); tc::mma_commit(&sink); }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (tc::elect_one()) tc::mma_commit(&sm.bar);
    __syncwarp();
    tc::mbar_wait(&sm.bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { cycles[blockIdx.x * 2] = t1 - t0; cycles[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

static void runpat(const char* name, int use_fence, int use_commit, int nS = 8, int nG = 4) {
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * 2 * 148);
  const size_t smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(kpat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 50;
  kpat<<<148, 128, smem>>>(4, use_fence, use_commit, nS, nG, cyc);
  kpat<<<148, 128, smem>>>(tiles, use_fence, use_commit, nS, nG, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[2 * 148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < 148; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  printf("%-52s issue %7.1f  complete %7.1f cyc/tile\n", name, issue / 148 / tiles, total / 148 / tiles);
  cudaFree(cyc);
}

static void run(const char* name, int mode, int N, int nacc, int hammer = 0, int iters = 240) {
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * 2 * 148);
  const size_t smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<148, hammer ? 288 : 128, smem>>>(mode, N, nacc, 12, cyc, hammer);
  k<<<148, hammer ? 288 : 128, smem>>>(mode, N, nacc, iters, cyc, hammer);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[2 * 148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < 148; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  issue /= 148; total /= 148;
  const double ideal = 128.0 * N / 256.0;
  printf("%-44s N=%3d acc=%d  issue %7.1f  complete %7.1f cyc/MMA   (ideal %5.1f -> %4.0f%% of peak)\n", name, N, nacc, issue / iters,
         total / iters, ideal, 100.0 * ideal / (total / iters));
  cudaFree(cyc);
}

int main() {
  for (int N : {16, 32, 64, 128, 256}) run("SS  K-major B, same accumulator", 0, N, 1);
  for (int N : {64, 128}) run("SS  K-major B, 2 accumulators", 0, N, 2);
  for (int N : {64, 128}) run("SS  K-major B, 3 accumulators", 0, N, 3);
  for (int N : {16, 32, 64, 112, 128}) run("TS  K-major B, same accumulator", 1, N, 1);
  for (int N : {16, 32, 64, 112, 128, 256}) run("TS  MN-major B, same accumulator", 2, N, 1);
  for (int N : {16, 32, 112}) run("TS  MN-major B, 2 accumulators", 2, N, 2);
  for (int N : {64, 128, 256}) run("SS  MN-major B, same accumulator", 3, N, 1);
  for (int N : {64, 112}) run("TS  K-major B + 8 warps tcgen05.ld", 1, N, 1, 1);
  for (int N : {64, 112}) run("TS  K-major B + 8 warps ld+ex2", 1, N, 1, 2);
  for (int N : {64, 128}) run("SS  K-major B + 8 warps ld+ex2", 0, N, 1, 2);
  runpat("dq pattern 8xS(N=64) + 4xG(N=112): no fence/commit", 0, 0);
  runpat("dq pattern: + commits", 0, 1);
  runpat("dq pattern: + fences", 1, 0);
  runpat("dq pattern: + fences + commits", 1, 1);
  runpat2<0, 8, 4>("unrolled 8+4: elect per block + syncwarp");
  runpat2<1, 8, 4>("unrolled 8+4: leader flag, no syncwarp");
  runpat2<2, 8, 4>("unrolled 8+4: lane 0 only");
  runpat2<3, 8, 4>("unrolled 8+4: leader + satisfied mbar wait + fence");
  runpat2<4, 8, 4>("unrolled 8+4: leader + satisfied mbar wait, no fence");
  runpat2<3, 8, 4>("8+4 wait+fence | 8 warps: ld", 1);
  runpat2<3, 8, 4>("8+4 wait+fence | 8 warps: ex2/fmul/cvt", 2);
  runpat2<3, 8, 4>("8+4 wait+fence | 8 warps: st", 4);
  runpat2<3, 8, 4>("8+4 wait+fence | 8 warps: ld+math+st", 7);
  runpat2<1, 8, 0>("unrolled 8+0: leader flag");
  runpat2<1, 0, 4>("unrolled 0+4: leader flag");
  runpat("only 8xS", 0, 0, 8, 0);
  runpat("only 4xG", 0, 0, 0, 4);
  return 0;
}
