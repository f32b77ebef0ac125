// Generic fp32 FFMA GEMM with functor-defined operands ("implicit GEMM" on CUDA cores).
//
//   C(m,n) = sum_{k in split} A(m,k) * B(k,n)        m<M, n<N, k<K
//
// A "problem" type P supplies the index maps, so one kernel serves every dense contraction of the
// fp32 path (qkv 1x1 projection, 3x3 strided conv fprop/dgrad/wgrad, out_proj, the key_rel weight
// reductions).  Split-K (gridDim.z) writes per-split partials that a second kernel sums in a fixed
// order, which keeps the fp32 path deterministic (no float atomics).
//
// P must provide:
//   int M, N, K, k_chunk;                                   // k_chunk = K range per split
//   static constexpr bool A_K_CONTIG, B_K_CONTIG;           // which index is contiguous in memory
//   struct ARow; struct BCol;                               // per-row / per-column decoded context
//   __device__ ARow a_row(int m) const;  __device__ float a(const ARow&, int k) const;
//   __device__ BCol b_col(int n) const;  __device__ float b(const BCol&, int k) const;
//   __device__ void store(int m, int n, float v, int split) const;
#pragma once
#include "common.cuh"

namespace aaconv {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <class P>
__global__ void __launch_bounds__(SG_THREADS) simt_gemm_kernel(const P p) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int split = blockIdx.z;
  const int kbeg = split * p.k_chunk;
  const int kend = min(p.K, kbeg + p.k_chunk);

  // thread -> (row, k) assignment for the global loads; 4 elements per thread per operand
  constexpr int NA = P::A_K_CONTIG ? 4 : 1;
  constexpr int NB = P::B_K_CONTIG ? 4 : 1;
  typename P::ARow arow[NA];
  typename P::BCol bcol[NB];
  bool aok[NA], bok[NB];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const int m = m0 + (P::A_K_CONTIG ? (tid >> 4) + 16 * i : (tid & 63));
    aok[i] = m < p.M;
    arow[i] = p.a_row(aok[i] ? m : 0);
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int n = n0 + (P::B_K_CONTIG ? (tid >> 4) + 16 * i : (tid & 63));
    bok[i] = n < p.N;
    bcol[i] = p.b_col(bok[i] ? n : 0);
  }

  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ka = k0 + (P::A_K_CONTIG ? (tid & 15) : (tid >> 6) + 4 * i);
      const int ia = P::A_K_CONTIG ? i : 0;
      ra[i] = (aok[ia] && ka < kend) ? p.a(arow[ia], ka) : 0.f;
      const int kb = k0 + (P::B_K_CONTIG ? (tid & 15) : (tid >> 6) + 4 * i);
      const int ib = P::B_K_CONTIG ? i : 0;
      rb[i] = (bok[ib] && kb < kend) ? p.b(bcol[ib], kb) : 0.f;
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (P::A_K_CONTIG) As[tid & 15][(tid >> 4) + 16 * i] = ra[i];
      else               As[(tid >> 6) + 4 * i][tid & 63] = ra[i];
      if (P::B_K_CONTIG) Bs[tid & 15][(tid >> 4) + 16 * i] = rb[i];
      else               Bs[(tid >> 6) + 4 * i][tid & 63] = rb[i];
    }
    __syncthreads();
    if (k0 + SG_BK < kend) fetch(k0 + SG_BK);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tx * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty * 4 + j;
      if (n < p.N) p.store(m, n, acc[i][j], split);
    }
  }
}

// out[i] (= or +=) sum_s partial[s*count + i]   -- fixed summation order
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                     int count, int splits, int accumulate);

template <class P>
int launch_simt_gemm(const P& p, int splits, cudaStream_t st, const char* name) {
  if (p.M <= 0 || p.N <= 0) return 0;
  dim3 grid(cdiv(p.M, SG_BM), cdiv(p.N, SG_BN), splits);
  simt_gemm_kernel<P><<<grid, SG_THREADS, 0, AACONV_ST(st)>>>(p);
  AACONV_LAUNCH_OK(name);
  return 0;
}

// Choose a split count so that a small-output / long-K GEMM fills the machine (148 SMs).
inline int pick_splits(int M, int N, int K) {
  const int tiles = cdiv(M, SG_BM) * cdiv(N, SG_BN);
  int s = cdiv(148 * 4, tiles);
  const int maxs = cdiv(K, 4 * SG_BK);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return s;
}
inline int chunk_for(int K, int splits) {
  int c = cdiv(K, splits);
  return cdiv(c, SG_BK) * SG_BK;
}

}  // namespace aaconv
