// Internal interface of the bf16 tcgen05 path.
#pragma once
#include "common.cuh"

namespace aaconv {
size_t bf16_saved_bytes(const Dims& d);
size_t bf16_scratch_bytes(const Dims& d);
int64_t bf16_saved_offset(const Dims& d, const char* name);
int bf16_forward(const Dims& d, const float* x, const aaconv_params* p, float* y, float* weights, void* saved,
                 void* scratch, cudaStream_t st);
int bf16_backward(const Dims& d, const float* x, const aaconv_params* p, const float* dy, const void* saved,
                  void* scratch, float* dx, const aaconv_param_grads* g, cudaStream_t st);
}  // namespace aaconv
