// Transition prologue fused into the AAConv2d boundary: the reference runs
//     InstanceNorm2d(affine=False, eps=1e-5) -> ReLU(inplace) -> AAConv2d          (models/attn_aug_conv.py:438-440)
// as three modules, i.e. two extra full passes over x forward (plus the layout / precision pack of the GEMM operand) and three
// backward.  Here:
//   in_stats          one block per (b, c) plane: mean and rstd = 1/sqrt(biased var + eps), two passes over a plane that
//                     stays in L1/L2 between them (no E[x^2] - E[x]^2 cancellation: the fp32 mode holds rtol 1e-4)
//   pack_x            x (fp32 | bf16, NCHW) -> xh (bf16, NHWC, channels padded to 64): the tcgen05 implicit-GEMM operand, with
//                     relu((x - mean) * rstd) applied on the way when fused -- the normalised tensor never exists in NCHW
//   in_relu_apply     fp32 NCHW materialisation of relu((x - mean) * rstd) (fp32 mode and the FFMA fallbacks only)
//   in_relu_bwd       one block per plane: g = dL/d(AAConv input), mask = x^ > 0,
//                     dx = rstd * (g.mask - mean(g.mask) - x^ * mean(g.mask * x^))           (adjoint of IN o ReLU, fixed-order sums)
// All reductions are per plane in a fixed order: bit-reproducible, no atomics.
#include <cuda_bf16.h>
#include "bf16_path.cuh"

namespace aaconv {

typedef __nv_bfloat16 bf16;

namespace {

template <class T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <class T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

// four consecutive elements (16-byte / 8-byte aligned)
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// fixed-order block sum of two values (256 threads): warp shuffles, then warp 0 over the 8 partials
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                     // sh may still be read from a previous call
  if (lane == 0) sh[warp] = make_float2(a, b);
  __syncthreads();
  float2 t = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < 8; ++w) { t.x += sh[w].x; t.y += sh[w].y; }
  return t;
}

template <class T>
__global__ void __launch_bounds__(256) in_stats_kernel(const T* __restrict__ x, float2* __restrict__ stats, int HW, float eps) {
  __shared__ float2 sh[8];
  const T* px = x + (size_t)blockIdx.x * HW;
  const bool vec = (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(px) & 15) == 0;
  float s = 0.f;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) { const float4 v = ld4(px + i); s += (v.x + v.y) + (v.z + v.w); }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) s += ldf(px + i);
  }
  const float mean = block_sum2(s, 0.f, sh).x / (float)HW;
  float q = 0.f;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) {
      const float4 v = ld4(px + i);
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) { const float a = ldf(px + i) - mean; q += a * a; }
  }
  const float var = block_sum2(q, 0.f, sh).x / (float)HW;
  if (threadIdx.x == 0) stats[blockIdx.x] = make_float2(mean, rsqrtf(var + eps));
}

// (B, C, HW) fp32|bf16 -> (B, HW, Cp) bf16, channels [C, Cp) zero; optional relu((x - mean) * rstd).  Tile = 64 channels x 64
// pixels: coalesced reads along pixels (4 per lane), 128 B coalesced writes along channels (4 x bf16 per lane).  HW % 4 == 0.
template <class T, bool FUSE>
__global__ void __launch_bounds__(256) pack_x_v4_kernel(const T* __restrict__ in, const float2* __restrict__ stats, bf16* __restrict__ out,
                                                        int C, int Cp, int HW) {
  __shared__ float t[64][65];                                   // [pixel][channel]
  const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const T* src = in + (size_t)b * C * HW;
  bf16* dst = out + (size_t)b * HW * Cp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int px = p0 + (lane & 15) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cl = warp * 8 + i * 2 + (lane >> 4), c = c0 + cl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C && px < HW) {
        v = ld4(src + (size_t)c * HW + px);                     // HW % 4 == 0: all-or-nothing
        if (FUSE) {
          const float2 st = __ldg(stats + (size_t)b * C + c);
          v.x = fmaxf((v.x - st.x) * st.y, 0.f); v.y = fmaxf((v.y - st.x) * st.y, 0.f);
          v.z = fmaxf((v.z - st.x) * st.y, 0.f); v.w = fmaxf((v.w - st.x) * st.y, 0.f);
        }
      }
      const int pl = (lane & 15) * 4;
      t[pl][cl] = v.x; t[pl + 1][cl] = v.y; t[pl + 2][cl] = v.z; t[pl + 3][cl] = v.w;
    }
  }
  __syncthreads();
  {
    const int cl = (lane & 15) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pl = warp * 8 + i * 2 + (lane >> 4), px = p0 + pl;
      if (px < HW && c0 + cl < Cp) st4(dst + (size_t)px * Cp + c0 + cl, make_float4(t[pl][cl], t[pl][cl + 1], t[pl][cl + 2], t[pl][cl + 3]));
    }
  }
}

// generic geometry: 32 x 32 tiles, scalar accesses
template <class T, bool FUSE>
__global__ void pack_x_kernel(const T* __restrict__ in, const float2* __restrict__ stats, bf16* __restrict__ out, int C, int Cp, int HW) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const T* src = in + (size_t)b * C * HW;
  bf16* dst = out + (size_t)b * HW * Cp;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, px = p0 + threadIdx.x;
    float v = 0.f;
    if (c < C && px < HW) {
      v = ldf(src + (size_t)c * HW + px);
      if (FUSE) { const float2 st = __ldg(stats + (size_t)b * C + c); v = fmaxf((v - st.x) * st.y, 0.f); }
    }
    t[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int px = p0 + i, c = c0 + threadIdx.x;
    if (px < HW && c < Cp) dst[(size_t)px * Cp + c] = __float2bfloat16(t[threadIdx.x][i]);
  }
}

template <class T>
__global__ void __launch_bounds__(256) in_relu_apply_kernel(const T* __restrict__ x, const float2* __restrict__ stats, float* __restrict__ out,
                                                            int HW, int fuse) {
  const size_t base = (size_t)blockIdx.x * HW;
  float2 st = make_float2(0.f, 1.f);
  if (fuse) st = stats[blockIdx.x];
  for (int i = threadIdx.x; i < HW; i += 256) {
    const float v = (ldf(x + base + i) - st.x) * st.y;
    out[base + i] = fuse ? fmaxf(v, 0.f) : v;
  }
}

// one block per (b, c) plane.  fuse = 0: plain copy / type conversion of g into dx.
template <class TX, class TG>
__global__ void __launch_bounds__(256) in_relu_bwd_kernel(const TX* __restrict__ x, const TG* __restrict__ g, const float2* __restrict__ stats,
                                                          TX* __restrict__ dx, int HW, int fuse) {
  __shared__ float2 sh[8];
  const size_t base = (size_t)blockIdx.x * HW;
  const TX* px = x + base;
  const TG* pg = g + base;
  TX* pd = dx + base;
  const bool vec = (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(px) | reinterpret_cast<uintptr_t>(pg) | reinterpret_cast<uintptr_t>(pd)) & 15) == 0;
  if (!fuse) {
    if (vec) for (int i = threadIdx.x * 4; i < HW; i += 1024) st4(pd + i, ld4(pg + i));
    else for (int i = threadIdx.x; i < HW; i += 256) stf(pd + i, ldf(pg + i));
    return;
  }
  const float2 st = stats[blockIdx.x];
  float s1 = 0.f, s2 = 0.f;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) {
      const float4 xv = ld4(px + i), gv = ld4(pg + i);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xh = (xs[u] - st.x) * st.y;
        const float gm = xh > 0.f ? gs[u] : 0.f;
        s1 += gm;
        s2 = fmaf(gm, xh, s2);
      }
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float xh = (ldf(px + i) - st.x) * st.y;
      const float gm = xh > 0.f ? ldf(pg + i) : 0.f;
      s1 += gm;
      s2 = fmaf(gm, xh, s2);
    }
  }
  const float2 tot = block_sum2(s1, s2, sh);
  const float m1 = tot.x / (float)HW, m2 = tot.y / (float)HW;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) {
      const float4 xv = ld4(px + i), gv = ld4(pg + i);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
      float o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xh = (xs[u] - st.x) * st.y;
        const float gm = xh > 0.f ? gs[u] : 0.f;
        o[u] = st.y * (gm - m1 - xh * m2);
      }
      st4(pd + i, make_float4(o[0], o[1], o[2], o[3]));
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float xh = (ldf(px + i) - st.x) * st.y;
      const float gm = xh > 0.f ? ldf(pg + i) : 0.f;
      stf(pd + i, st.y * (gm - m1 - xh * m2));
    }
  }
}

}  // namespace

int in_stats(const Dims& d, const void* x, float* stats, cudaStream_t st) {
  const int planes = d.B * d.Cin, HW = d.Hin * d.Win;
  if (d.x_bf16) in_stats_kernel<bf16><<<planes, 256, 0, AACONV_ST(st)>>>(static_cast<const bf16*>(x), reinterpret_cast<float2*>(stats), HW, d.in_eps);
  else in_stats_kernel<float><<<planes, 256, 0, AACONV_ST(st)>>>(static_cast<const float*>(x), reinterpret_cast<float2*>(stats), HW, d.in_eps);
  AACONV_LAUNCH_OK("in_stats");
  return 0;
}

template <class T>
static int pack_x_t(const T* in, const float* stats, void* out, int B, int C, int Cp, int HW, cudaStream_t st) {
  const float2* s2 = reinterpret_cast<const float2*>(stats);
  bf16* o = static_cast<bf16*>(out);
  if (HW % 4 == 0 && Cp % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    dim3 grid(cdiv(HW, 64), cdiv(Cp, 64), B);
    if (stats) pack_x_v4_kernel<T, true><<<grid, 256, 0, AACONV_ST(st)>>>(in, s2, o, C, Cp, HW);
    else pack_x_v4_kernel<T, false><<<grid, 256, 0, AACONV_ST(st)>>>(in, s2, o, C, Cp, HW);
  } else {
    dim3 grid(cdiv(HW, 32), cdiv(Cp, 32), B), block(32, 8);
    if (stats) pack_x_kernel<T, true><<<grid, block, 0, AACONV_ST(st)>>>(in, s2, o, C, Cp, HW);
    else pack_x_kernel<T, false><<<grid, block, 0, AACONV_ST(st)>>>(in, s2, o, C, Cp, HW);
  }
  AACONV_LAUNCH_OK(stats ? "pack_nhwc_bf16_in_relu" : "pack_nhwc_bf16");
  return 0;
}

// stats == NULL: plain layout / precision pack
int pack_x(const void* in, int in_bf16, const float* stats, void* out, int B, int C, int Cp, int HW, cudaStream_t st) {
  return in_bf16 ? pack_x_t(static_cast<const bf16*>(in), stats, out, B, C, Cp, HW, st)
                 : pack_x_t(static_cast<const float*>(in), stats, out, B, C, Cp, HW, st);
}
int pack_nhwc_bf16(const float* in, void* out, int B, int C, int Cp, int HW, cudaStream_t st) {
  return pack_x_t(in, nullptr, out, B, C, Cp, HW, st);
}

int in_relu_apply(const Dims& d, const void* x, const float* stats, float* out, cudaStream_t st) {
  const int planes = d.B * d.Cin, HW = d.Hin * d.Win;
  const float2* s2 = reinterpret_cast<const float2*>(stats);
  if (d.x_bf16) in_relu_apply_kernel<bf16><<<planes, 256, 0, AACONV_ST(st)>>>(static_cast<const bf16*>(x), s2, out, HW, d.fuse_in);
  else in_relu_apply_kernel<float><<<planes, 256, 0, AACONV_ST(st)>>>(static_cast<const float*>(x), s2, out, HW, d.fuse_in);
  AACONV_LAUNCH_OK("in_relu_apply");
  return 0;
}

// g: gradient with respect to the AAConv2d input (fp32, or bf16 when g_bf16), dense NCHW.  dx in the type of x.
int in_relu_bwd(const Dims& d, const void* x, const void* g, int g_bf16, const float* stats, void* dx, cudaStream_t st) {
  const int planes = d.B * d.Cin, HW = d.Hin * d.Win;
  const float2* s2 = reinterpret_cast<const float2*>(stats);
  if (d.x_bf16) {
    const bf16* xx = static_cast<const bf16*>(x);
    bf16* dd = static_cast<bf16*>(dx);
    if (g_bf16) in_relu_bwd_kernel<bf16, bf16><<<planes, 256, 0, AACONV_ST(st)>>>(xx, static_cast<const bf16*>(g), s2, dd, HW, d.fuse_in);
    else in_relu_bwd_kernel<bf16, float><<<planes, 256, 0, AACONV_ST(st)>>>(xx, static_cast<const float*>(g), s2, dd, HW, d.fuse_in);
  } else {
    const float* xx = static_cast<const float*>(x);
    float* dd = static_cast<float*>(dx);
    if (g_bf16) in_relu_bwd_kernel<float, bf16><<<planes, 256, 0, AACONV_ST(st)>>>(xx, static_cast<const bf16*>(g), s2, dd, HW, d.fuse_in);
    else in_relu_bwd_kernel<float, float><<<planes, 256, 0, AACONV_ST(st)>>>(xx, static_cast<const float*>(g), s2, dd, HW, d.fuse_in);
  }
  AACONV_LAUNCH_OK(d.fuse_in ? "in_relu_bwd" : "dx_convert");
  return 0;
}

}  // namespace aaconv
