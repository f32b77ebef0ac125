// Internal interface of the fp32 (FFMA) path.  Layouts: q,k (B,nh,L,dkh) with q pre-scaled by dkh^-0.5;
// v,o,dO (B,nh,L,dvh); lse,delta (B,nh,L); rw/drw (B,nh,L,2W-1); rh/drh (B,nh,L,2H-1).
#pragma once
#include "common.cuh"

namespace aaconv {

// fp32_gemms.cu
int f32_conv_fwd(const Dims& d, const float* x, const float* w, void* y, cudaStream_t st);   // y: d.y_bf16 / d.y_bs
int f32_qkv_fwd(const Dims& d, const float* x, const float* w, float* q, float* k, float* v, cudaStream_t st);
int f32_out_fwd(const Dims& d, const float* o, const float* w, void* y, cudaStream_t st);
int f32_out_bwd(const Dims& d, const float* dy, const float* o, const float* w, float* d_o, float* dw,
                float* partial, cudaStream_t st);
int f32_out_bwd_weight(const Dims& d, const float* dy, const float* o, float* dw, float* partial, cudaStream_t st);
int f32_qkv_bwd(const Dims& d, const float* x, const float* w, const float* dq, const float* dk,
                const float* dv, float* dw, float* dx, int dx_accumulate, float* partial, cudaStream_t st);
int f32_conv_bwd(const Dims& d, const float* x, const float* w, const float* dy, float* dx, float* dw,
                 float* partial, cudaStream_t st);
int f32_rel_weight_grad(const Dims& d, const float* q, const float* dr, int R, float* dkr, float* partial,
                        cudaStream_t st);
size_t f32_partial_floats(const Dims& d);

// fp32_attn.cu
int f32_rel_fwd(const Dims& d, const float* q, const float* krw, const float* krh, float* rw, float* rh,
                cudaStream_t st);
int f32_attn_fwd(const Dims& d, const float* q, const float* k, const float* v, const float* rw,
                 const float* rh, float* o, float* lse, cudaStream_t st);
int f32_attn_weights(const Dims& d, const float* q, const float* k, const float* rw, const float* rh,
                     const float* lse, float* weights, cudaStream_t st);
int f32_delta(const Dims& d, const float* d_o, const float* o, float* delta, cudaStream_t st);
int f32_attn_bwd(const Dims& d, const float* q, const float* k, const float* v, const float* rw,
                 const float* rh, const float* lse, const float* d_o, const float* delta, float* dq,
                 float* dk, float* dv, float* drw, float* drh, cudaStream_t st);
int f32_rel_bwd_dq(const Dims& d, const float* krw, const float* krh, const float* drw, const float* drh,
                   float* dq, cudaStream_t st);

// fp32_path.cu  (orchestration)
size_t f32_saved_bytes(const Dims& d);           // q, k, v, o, lse (the block the bf16 path shares)
size_t f32_saved_bytes_io(const Dims& d);        // + the InstanceNorm statistics of the fused prologue
size_t f32_scratch_bytes(const Dims& d);
int64_t f32_saved_offset(const Dims& d, const char* name);
int f32_forward(const Dims& d, const void* x, const aaconv_params* p, void* y, float* weights, void* saved,
                void* scratch, cudaStream_t st);
int f32_backward(const Dims& d, const void* x, const aaconv_params* p, const float* dy, void* saved,
                 void* scratch, void* dx, const aaconv_param_grads* g, cudaStream_t st);

}  // namespace aaconv
