// Orchestration of the bf16 tensor-core path.  Stages not yet moved to tcgen05 run the fp32 FFMA kernels
// (same layouts), so the path is always complete; DESIGN.md lists which stage runs where.
#include "bf16_path.cuh"
#include "fp32_path.cuh"

namespace aaconv {

size_t bf16_saved_bytes(const Dims& d) { return f32_saved_bytes(d); }
size_t bf16_scratch_bytes(const Dims& d) {
  if (tc_attn_supported(d)) return 0;
  return align256(f32_scratch_bytes(d)) + tc_attn_operand_bytes(d, nullptr, nullptr, nullptr);
}
int64_t bf16_saved_offset(const Dims& d, const char* name) { return f32_saved_offset(d, name); }

namespace {
template <class T>
T* at(void* base, int64_t off) { return reinterpret_cast<T*>(static_cast<char*>(base) + off); }
}  // namespace

int bf16_forward(const Dims& d, const float* x, const aaconv_params* p, float* y, float* weights, void* saved,
                 void* scratch, cudaStream_t st) {
  AACONV_TRY(tc_attn_supported(d));
  float* q = at<float>(saved, f32_saved_offset(d, "q"));
  float* k = at<float>(saved, f32_saved_offset(d, "k"));
  float* v = at<float>(saved, f32_saved_offset(d, "v"));
  float* o = at<float>(saved, f32_saved_offset(d, "o"));
  float* lse = at<float>(saved, f32_saved_offset(d, "lse"));
  void* operands = at<char>(scratch, (int64_t)align256(f32_scratch_bytes(d)));
  AACONV_TRY(f32_conv_fwd(d, x, p->conv_w, y, st));
  AACONV_TRY(f32_qkv_fwd(d, x, p->qkv_w, q, k, v, st));
  AACONV_TRY(tc_attn_fwd(d, q, k, v, p->key_rel_w, p->key_rel_h, operands, o, lse, st));
  if (weights) {   // visualise path only: fp32 map from the saved statistics
    float* rw = at<float>(scratch, 0);
    float* rh = rw + align256((size_t)d.BN * d.L * d.RW * sizeof(float)) / sizeof(float);
    AACONV_TRY(f32_rel_fwd(d, q, p->key_rel_w, p->key_rel_h, rw, rh, st));
    AACONV_TRY(f32_attn_weights(d, q, k, rw, rh, lse, weights, st));
  }
  AACONV_TRY(f32_out_fwd(d, o, p->out_w, y, st));
  return 0;
}

int bf16_backward(const Dims& d, const float* x, const aaconv_params* p, const float* dy, const void* saved,
                  void* scratch, float* dx, const aaconv_param_grads* g, cudaStream_t st) {
  return f32_backward(d, x, p, dy, saved, scratch, dx, g, st);
}

}  // namespace aaconv
