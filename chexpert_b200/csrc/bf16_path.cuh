// Internal interface of the bf16 tcgen05 path.
#pragma once
#include "common.cuh"

namespace aaconv {
// attn_tc.cu
int tc_attn_supported(const Dims& d);
size_t tc_attn_operand_bytes(const Dims& d, size_t* qa_off, size_t* ka_off, size_t* vt_off);
int tc_attn_fwd(const Dims& d, const float* q, const float* k, const float* v, const float* krw, const float* krh,
                void* operands, float* o, float* lse, cudaStream_t st);

// bf16_path.cu
size_t bf16_saved_bytes(const Dims& d);
size_t bf16_scratch_bytes(const Dims& d);
int64_t bf16_saved_offset(const Dims& d, const char* name);
int bf16_forward(const Dims& d, const float* x, const aaconv_params* p, float* y, float* weights, void* saved,
                 void* scratch, cudaStream_t st);
int bf16_backward(const Dims& d, const float* x, const aaconv_params* p, const float* dy, const void* saved,
                  void* scratch, float* dx, const aaconv_param_grads* g, cudaStream_t st);
}  // namespace aaconv
