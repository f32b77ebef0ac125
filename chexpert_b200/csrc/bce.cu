// Multi-label BCE-with-logits on the competition classes (chexpert.py:530,160,205; dataset.py:25,139,142).
// One launch produces the element losses (eval path), the scalar train loss (sum over classes, mean over
// batch) and d loss / d z.  (B x C) is tiny (16 x 5 in training), so this is a latency item: one block,
// warp-shuffle reduction, no atomics.
#include "common.cuh"

namespace aaconv {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ z, const float* __restrict__ targets,
                                                  int ld, const int32_t* __restrict__ cols, int B, int C,
                                                  float* __restrict__ el, float* __restrict__ loss,
                                                  float* __restrict__ dz, const float* __restrict__ grad_scale) {
  __shared__ float warp_part[8];
  const float gs = (grad_scale ? grad_scale[0] : 1.f) / (float)B;
  float part = 0.f;
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const int b = i / C, c = i - b * C;
    float t = cols ? targets[(size_t)b * ld + cols[c]] : targets[(size_t)b * ld + c];
    if (cols) {                 // U-Ones policy on raw labels: blank (NaN) -> 0, uncertain (-1) -> 1
      if (t != t) t = 0.f;
      else if (t == -1.f) t = 1.f;
    }
    const float zi = z[i];
    const float e = fmaxf(zi, 0.f) - zi * t + log1pf(expf(-fabsf(zi)));
    if (el) el[i] = e;
    if (dz) dz[i] = (1.f / (1.f + expf(-zi)) - t) * gs;
    part += e;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? warp_part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && loss) loss[0] = v / (float)B;
  }
}

int bce_launch(const float* z, const float* targets, int ld, const int32_t* cols, int B, int C, float* el,
               float* loss, float* dz, const float* grad_scale, cudaStream_t st) {
  bce_kernel<<<1, 256, 0, AACONV_ST(st)>>>(z, targets, ld, cols, B, C, el, loss, dz, grad_scale);
  AACONV_LAUNCH_OK("bce");
  return 0;
}

}  // namespace aaconv
