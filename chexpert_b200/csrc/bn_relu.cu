// Fused BatchNorm2d (training mode, batch statistics) + ReLU over a channel slice of the pre-allocated DenseBlock feature buffer
// (the reference runs torchvision's _DenseLayer: norm1 -> relu1 -> conv1 -> norm2 -> relu2 -> conv2,
// tv densenet.py:31-95 under models/attn_aug_conv.py:479-482).
//
// Why it exists: layer i of a dense block normalises channels [0, c_i) of ALL features so far.  With the feature buffer that
// input is a strided view (batch stride = C_total * H * W), for which PyTorch's batch_norm falls off its fast path (a generic
// Welford reduce_kernel + three elementwise kernels: the buffered block was 7 % slower than torch.cat), and even on the fast
// path BN + ReLU forward / backward are 4 + 8 passes over the activations with ONE CTA per channel for the reductions (64
// CTAs on 148 SMs at the first layers).  Here:
//   forward   bn_stats   grid (C, B): per-(b, c)-plane (count, mean, M2) partials, Chan-merged later -> no cancellation.
//                        In a dense block layer i normalises channels [0, c_i) of which [0, c_{i-1}) were already reduced by
//                        layer i-1 (same data, same statistics): the caller keeps ONE partial buffer per block and passes
//                        `stats_valid_channels`, so only the layer's 32 new channels are reduced (the pass shrinks ~5x)
//             bn_apply   grid (C, B): merges the B partials of its channel (fixed order), y = relu(w (x - mean) rstd + b);
//                        the b = 0 block also writes mean / rstd for backward and updates running_mean / running_var
//   backward  bn_bwd_red grid (C, B): g = dy * [y > 0], partial (sum g, sum g x^) per plane
//             bn_bwd_dx  grid (C, B): merges the partials (fixed order), dx = w rstd (g - mean(g) - x^ mean(g x^));
//                        the b = 0 block writes dweight = sum g x^ and dbias = sum g
// 3 / 5 passes, all reductions in a fixed order (bit-reproducible), x read through its strides, y / dx dense.
#include <cuda_bf16.h>
#include "common.cuh"

namespace aaconv {

typedef __nv_bfloat16 bf16;

namespace {

template <class T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld1<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <class T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }
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

// fixed-order block sum of two values (256 threads)
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = make_float2(a, b);
  __syncthreads();
  float2 t = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < 8; ++w) { t.x += sh[w].x; t.y += sh[w].y; }
  return t;
}

// partial[(c * B + b)] = (mean, M2) of plane (b, c): two passes over a plane that stays in L1 / L2 between them
template <class T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, long long xbs, int HW, float2* __restrict__ partial, int c_from) {
  __shared__ float2 sh[8];
  const int c = c_from + blockIdx.x, b = blockIdx.y, B = gridDim.y;
  const T* px = x + (size_t)b * xbs + (size_t)c * HW;
  const bool vec = (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(px) & 15) == 0;
  float s = 0.f;
  if (vec) for (int i = threadIdx.x * 4; i < HW; i += 1024) { const float4 v = ld4(px + i); s += (v.x + v.y) + (v.z + v.w); }
  else for (int i = threadIdx.x; i < HW; i += 256) s += ld1(px + i);
  const float mean = block_sum2(s, 0.f, sh).x / (float)HW;
  float q = 0.f;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) {
      const float4 v = ld4(px + i);
      const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) { const float a0 = ld1(px + i) - mean; q += a0 * a0; }
  }
  const float m2 = block_sum2(q, 0.f, sh).x;
  if (threadIdx.x == 0) partial[(size_t)c * B + b] = make_float2(mean, m2);
}

// Chan merge of the B plane partials of channel c, in batch order: -> (mean, biased var).  The partials are fetched by the
// block's threads in parallel (ONE memory latency, not B dependent ones: with the merge as a serial load loop every block
// spent ~5 us before touching its plane and bn_apply averaged 29 us), then combined by thread 0 in a fixed order.
__device__ __forceinline__ float2 merge_stats(const float2* __restrict__ partial, int c, int B, int HW, float2* sh /*[257]*/) {
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int b0 = 0; b0 < B; b0 += 256) {
    const int nb_ = min(256, B - b0);
    __syncthreads();
    if ((int)threadIdx.x < nb_) sh[threadIdx.x] = partial[(size_t)c * B + b0 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 0; i < nb_; ++i) {
        const float2 p = sh[i];
        const float nb = (float)HW, tot = n + nb, d = p.x - mean;
        mean += d * (nb / tot);
        m2 += p.y + d * d * (n * nb / tot);
        n = tot;
      }
    }
  }
  if (threadIdx.x == 0) sh[256] = make_float2(mean, m2 / n);
  __syncthreads();
  return sh[256];
}
// plain sums of the B plane partials of channel c, in batch order
__device__ __forceinline__ float2 merge_sums(const float2* __restrict__ partial, int c, int B, float2* sh /*[257]*/) {
  float s1 = 0.f, s2 = 0.f;
  for (int b0 = 0; b0 < B; b0 += 256) {
    const int nb_ = min(256, B - b0);
    __syncthreads();
    if ((int)threadIdx.x < nb_) sh[threadIdx.x] = partial[(size_t)c * B + b0 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < nb_; ++i) { s1 += sh[i].x; s2 += sh[i].y; }
  }
  if (threadIdx.x == 0) sh[256] = make_float2(s1, s2);
  __syncthreads();
  return sh[256];
}

template <class T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, long long xbs, int HW, const float2* __restrict__ partial,
                                                       const float* __restrict__ weight, const float* __restrict__ bias,
                                                       float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                                       float eps, T* __restrict__ y, float2* __restrict__ saved) {
  __shared__ float2 shm[257];
  const int c = blockIdx.x, b = blockIdx.y, B = gridDim.y, C = gridDim.x;
  const float2 st = merge_stats(partial, c, B, HW, shm);
  const float rstd = rsqrtf(st.y + eps);
  if (b == 0 && threadIdx.x == 0) {
    saved[c] = make_float2(st.x, rstd);
    if (running_mean) {                                    // nn.BatchNorm2d: unbiased variance in the running estimate
      const float n = (float)B * HW;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * st.x;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * st.y * (n / fmaxf(n - 1.f, 1.f));
    }
  }
  const float sc = weight[c] * rstd, sh = bias[c] - st.x * sc;
  const T* px = x + (size_t)b * xbs + (size_t)c * HW;
  T* py = y + ((size_t)b * C + c) * HW;
  const bool vec = (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(px) | reinterpret_cast<uintptr_t>(py)) & 15) == 0;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) {
      const float4 v = ld4(px + i);
      st4(py + i, make_float4(fmaxf(fmaf(v.x, sc, sh), 0.f), fmaxf(fmaf(v.y, sc, sh), 0.f), fmaxf(fmaf(v.z, sc, sh), 0.f),
                              fmaxf(fmaf(v.w, sc, sh), 0.f)));
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) st1(py + i, fmaxf(fmaf(ld1(px + i), sc, sh), 0.f));
  }
}

// g = dy where the ReLU output is positive (recomputed from x and the saved statistics: y itself is not kept)
template <class T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* __restrict__ x, long long xbs, int HW, const T* __restrict__ dy,
                                                            const float2* __restrict__ saved, const float* __restrict__ weight,
                                                            const float* __restrict__ bias, float2* __restrict__ partial) {
  __shared__ float2 sh[8];
  const int c = blockIdx.x, b = blockIdx.y, B = gridDim.y, C = gridDim.x;
  const float2 st = saved[c];
  const float w = weight[c], bb = bias[c];
  const T* px = x + (size_t)b * xbs + (size_t)c * HW;
  const T* pg = dy + ((size_t)b * C + c) * HW;
  const bool vec = (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(px) | reinterpret_cast<uintptr_t>(pg)) & 15) == 0;
  float s1 = 0.f, s2 = 0.f;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) {
      const float4 xv = ld4(px + i), gv = ld4(pg + i);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xh = (xs[u] - st.x) * st.y;
        const float g = fmaf(w, xh, bb) > 0.f ? gs[u] : 0.f;
        s1 += g;
        s2 = fmaf(g, xh, s2);
      }
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float xh = (ld1(px + i) - st.x) * st.y;
      const float g = fmaf(w, xh, bb) > 0.f ? ld1(pg + i) : 0.f;
      s1 += g;
      s2 = fmaf(g, xh, s2);
    }
  }
  const float2 t = block_sum2(s1, s2, sh);
  if (threadIdx.x == 0) partial[(size_t)c * B + b] = t;
}

template <class T>
__global__ void __launch_bounds__(256) bn_bwd_dx_kernel(const T* __restrict__ x, long long xbs, int HW, const T* __restrict__ dy,
                                                        const float2* __restrict__ saved, const float* __restrict__ weight,
                                                        const float* __restrict__ bias, const float2* __restrict__ partial,
                                                        T* __restrict__ dx, float* __restrict__ dweight, float* __restrict__ dbias) {
  __shared__ float2 shm[257];
  const int c = blockIdx.x, b = blockIdx.y, B = gridDim.y, C = gridDim.x;
  const float2 tot = merge_sums(partial, c, B, shm);                     // batch order
  const float s1 = tot.x, s2 = tot.y;
  if (b == 0 && threadIdx.x == 0) {
    if (dweight) dweight[c] = s2;
    if (dbias) dbias[c] = s1;
  }
  if (!dx) return;
  const float2 st = saved[c];
  const float w = weight[c], bb = bias[c], inv_n = 1.f / ((float)B * HW);
  const float m1 = s1 * inv_n, m2 = s2 * inv_n, k = w * st.y;
  const T* px = x + (size_t)b * xbs + (size_t)c * HW;
  const T* pg = dy + ((size_t)b * C + c) * HW;
  T* pd = dx + ((size_t)b * C + c) * HW;
  const bool vec = (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(px) | reinterpret_cast<uintptr_t>(pg) | reinterpret_cast<uintptr_t>(pd)) & 15) == 0;
  if (vec) {
    for (int i = threadIdx.x * 4; i < HW; i += 1024) {
      const float4 xv = ld4(px + i), gv = ld4(pg + i);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
      float o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xh = (xs[u] - st.x) * st.y;
        const float g = fmaf(w, xh, bb) > 0.f ? gs[u] : 0.f;
        o[u] = k * (g - m1 - xh * m2);
      }
      st4(pd + i, make_float4(o[0], o[1], o[2], o[3]));
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float xh = (ld1(px + i) - st.x) * st.y;
      const float g = fmaf(w, xh, bb) > 0.f ? ld1(pg + i) : 0.f;
      st1(pd + i, k * (g - m1 - xh * m2));
    }
  }
}

template <class T>
int fwd_t(const T* x, long long xbs, int B, int C, int HW, const float* w, const float* b, float* rm, float* rv, float momentum, float eps,
          T* y, float* saved, float* ws, int c_from, cudaStream_t st) {
  dim3 grid(C, B);
  if (c_from < C) {                                       // plane statistics of channels [c_from, C); the rest is already in `ws`
    bn_stats_kernel<T><<<dim3(C - c_from, B), 256, 0, AACONV_ST(st)>>>(x, xbs, HW, reinterpret_cast<float2*>(ws), c_from);
    AACONV_LAUNCH_OK("bn_stats");
  }
  bn_apply_kernel<T><<<grid, 256, 0, AACONV_ST(st)>>>(x, xbs, HW, reinterpret_cast<const float2*>(ws), w, b, rm, rv, momentum, eps, y,
                                                      reinterpret_cast<float2*>(saved));
  AACONV_LAUNCH_OK("bn_relu_apply");
  return 0;
}
template <class T>
int bwd_t(const T* x, long long xbs, int B, int C, int HW, const T* dy, const float* saved, const float* w, const float* b, T* dx, float* dw,
          float* db, float* ws, cudaStream_t st) {
  dim3 grid(C, B);
  bn_bwd_reduce_kernel<T><<<grid, 256, 0, AACONV_ST(st)>>>(x, xbs, HW, dy, reinterpret_cast<const float2*>(saved), w, b,
                                                           reinterpret_cast<float2*>(ws));
  AACONV_LAUNCH_OK("bn_relu_bwd_reduce");
  bn_bwd_dx_kernel<T><<<grid, 256, 0, AACONV_ST(st)>>>(x, xbs, HW, dy, reinterpret_cast<const float2*>(saved), w, b,
                                                       reinterpret_cast<const float2*>(ws), dx, dw, db);
  AACONV_LAUNCH_OK("bn_relu_bwd_dx");
  return 0;
}

}  // namespace
}  // namespace aaconv

using namespace aaconv;

extern "C" {

size_t aaconv_bn_relu_workspace_bytes(int B, int C) { return (B > 0 && C > 0) ? align256(sizeof(float) * 2 * (size_t)B * C) : 0; }

int aaconv_bn_relu_forward(const void* x, int dtype, int B, int C, int HW, int64_t x_batch_stride, const float* weight, const float* bias,
                           float* running_mean, float* running_var, float momentum, float eps, void* y, float* saved, void* workspace,
                           int stats_valid_channels, void* stream) {
  if (!x || !weight || !bias || !y || !saved || !workspace || B <= 0 || C <= 0 || HW <= 0 || B > 65535 || stats_valid_channels < 0 ||
      stats_valid_channels > C)
    return fail(AACONV_E_ARG, "bad bn_relu_forward arguments");
  if ((dtype != AACONV_FP32 && dtype != AACONV_BF16) || x_batch_stride < (int64_t)C * HW) return fail(AACONV_E_ARG, "bad bn_relu dtype / stride");
  if ((running_mean == nullptr) != (running_var == nullptr)) return fail(AACONV_E_ARG, "running_mean and running_var come together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  return dtype == AACONV_BF16
             ? fwd_t(static_cast<const bf16*>(x), x_batch_stride, B, C, HW, weight, bias, running_mean, running_var, momentum, eps,
                     static_cast<bf16*>(y), saved, ws, stats_valid_channels, st)
             : fwd_t(static_cast<const float*>(x), x_batch_stride, B, C, HW, weight, bias, running_mean, running_var, momentum, eps,
                     static_cast<float*>(y), saved, ws, stats_valid_channels, st);
}

int aaconv_bn_relu_backward(const void* x, int dtype, int B, int C, int HW, int64_t x_batch_stride, const void* dy, const float* saved,
                            const float* weight, const float* bias, void* dx, float* dweight, float* dbias, void* workspace, void* stream) {
  if (!x || !dy || !saved || !weight || !bias || !workspace || B <= 0 || C <= 0 || HW <= 0 || B > 65535)
    return fail(AACONV_E_ARG, "bad bn_relu_backward arguments");
  if ((dtype != AACONV_FP32 && dtype != AACONV_BF16) || x_batch_stride < (int64_t)C * HW) return fail(AACONV_E_ARG, "bad bn_relu dtype / stride");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  return dtype == AACONV_BF16 ? bwd_t(static_cast<const bf16*>(x), x_batch_stride, B, C, HW, static_cast<const bf16*>(dy), saved, weight, bias,
                                      static_cast<bf16*>(dx), dweight, dbias, ws, st)
                              : bwd_t(static_cast<const float*>(x), x_batch_stride, B, C, HW, static_cast<const float*>(dy), saved, weight,
                                      bias, static_cast<float*>(dx), dweight, dbias, ws, st);
}

}  // extern "C"
