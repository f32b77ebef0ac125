// Ensemble mean of logits and per-class AUROC on the device.
//   chexpert.py:233      outputs = torch.stack(outputs, dim=2).mean(2)           -> ensemble_mean_kernel
//   chexpert.py:130-135  roc_curve + auc per class on the raw logits             -> auroc_count_kernel + auroc_final_kernel
// The area under sklearn's ROC polygon equals the tie-corrected Mann-Whitney statistic
//   AUROC = ( #{(i,j): t_i = 1, t_j = 0, z_i > z_j} + 0.5 #{... z_i == z_j} ) / (n_pos n_neg),
// counted here exactly in integers (order-independent, so the result is deterministic); a class with a single label
// value gives NaN, as sklearn does (the reference averages with np.nanmean, chexpert.py:189).
#include "common.cuh"

namespace aaconv {

__global__ void ensemble_mean_kernel(const float* __restrict__ logits, int M, size_t n, float* __restrict__ mean) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += logits[(size_t)m * n + i];   // fixed order
    mean[i] = s / (float)M;
  }
}

// counters[c] = { 2 * greater + equal, n_pos, n_neg };  grid (C, slices of the i range)
__global__ void __launch_bounds__(256) auroc_count_kernel(const float* __restrict__ z, const float* __restrict__ t, int N, int C,
                                                          unsigned long long* __restrict__ counters) {
  const int c = blockIdx.x;
  unsigned long long score = 0, npos = 0, nneg = 0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < N; i += gridDim.y * blockDim.x) {
    const float ti = t[(size_t)i * C + c];
    if (ti > 0.5f) {
      ++npos;
      const float zi = z[(size_t)i * C + c];
      for (int j = 0; j < N; ++j) {
        if (t[(size_t)j * C + c] > 0.5f) continue;
        const float zj = z[(size_t)j * C + c];
        score += zi > zj ? 2u : (zi == zj ? 1u : 0u);
      }
    } else {
      ++nneg;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    score += __shfl_xor_sync(0xffffffffu, score, o);
    npos += __shfl_xor_sync(0xffffffffu, npos, o);
    nneg += __shfl_xor_sync(0xffffffffu, nneg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counters[c * 3 + 0], score);
    atomicAdd(&counters[c * 3 + 1], npos);
    atomicAdd(&counters[c * 3 + 2], nneg);
  }
}

__global__ void auroc_final_kernel(const unsigned long long* __restrict__ counters, int C, float* __restrict__ auroc) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const double pairs = (double)counters[c * 3 + 1] * (double)counters[c * 3 + 2];
  auroc[c] = pairs > 0 ? (float)((double)counters[c * 3] / (2.0 * pairs)) : __int_as_float(0x7fc00000);
}

int ensemble_mean_launch(const float* logits, int M, size_t n, float* mean, cudaStream_t st) {
  const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  ensemble_mean_kernel<<<blocks, 256, 0, AACONV_ST(st)>>>(logits, M, n, mean);
  AACONV_LAUNCH_OK("ensemble_mean");
  return 0;
}

int auroc_launch(const float* z, const float* t, int N, int C, float* auroc, void* workspace, cudaStream_t st) {
  unsigned long long* counters = static_cast<unsigned long long*>(workspace);
  AACONV_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * 3 * C, st));
  const int slices = (N + 255) / 256 < 64 ? (N + 255) / 256 : 64;
  auroc_count_kernel<<<dim3(C, slices), 256, 0, AACONV_ST(st)>>>(z, t, N, C, counters);
  AACONV_LAUNCH_OK("auroc_count");
  auroc_final_kernel<<<1, ((C + 31) / 32) * 32, 0, AACONV_ST(st)>>>(counters, C, auroc);
  AACONV_LAUNCH_OK("auroc_final");
  return 0;
}

}  // namespace aaconv
