// Orchestration of the bf16 tensor-core path.  Stages not yet moved to tcgen05 run the fp32 FFMA kernels
// (same layouts), so the path is always complete; DESIGN.md lists which stage runs where.
#include <algorithm>
#include <cstring>
#include "bf16_path.cuh"
#include "fp32_path.cuh"

namespace aaconv {

namespace {
struct Scratch {
  float *d_o, *delta, *dq, *dk, *dv, *partial, *dqa, *rw, *rh, *o_tmp, *lse_tmp;
  void *qa, *ka, *fwd_operands;
  TcGemmBufs gemm;
  bool gemm_ok;
  size_t bytes;
  Scratch(const Dims& d, void* base, int want_weights) {
    Carver c(base);
    const size_t rows = (size_t)d.BN * d.L;
    const AugLayout a = aug_layout(d);
    d_o = c.take<float>(rows * d.dvh);
    delta = c.take<float>(rows);
    dq = c.take<float>(rows * d.dkh);
    dk = c.take<float>(rows * d.dkh);
    dv = c.take<float>(rows * d.dvh);
    partial = c.take<float>(std::max(f32_partial_floats(d), tc_gemm_supported(d) == 0 ? tc_wgrad_partial_floats(d) : 0));
    dqa = c.take<float>(rows * a.KD);
    qa = c.take<uint16_t>(rows * a.KP);
    ka = c.take<uint16_t>(rows * a.KP);
    fwd_operands = c.take<char>(tc_attn_operand_bytes(d, nullptr, nullptr, nullptr));
    gemm_ok = tc_gemm_supported(d) == 0;
    {
      char* gb = c.take<char>(gemm_ok ? tc_gemm_bufs(d, nullptr).bytes : 0);
      gemm = tc_gemm_bufs(d, gb);
    }
    const bool w = want_weights && d.relative;
    rw = c.take<float>(w ? rows * d.RW : 0);
    rh = c.take<float>(w ? rows * d.RH : 0);
    o_tmp = c.take<float>(want_weights ? rows * d.dvh : 0);
    lse_tmp = c.take<float>(want_weights ? rows : 0);
    bytes = c.off;
  }
};
template <class T>
T* at(const void* base, int64_t off) { return reinterpret_cast<T*>(static_cast<char*>(const_cast<void*>(base)) + off); }
}  // namespace

size_t bf16_saved_bytes(const Dims& d) { return f32_saved_bytes(d); }
size_t bf16_scratch_bytes(const Dims& d, int want_weights) {
  if (aug_supported(d) || tc_attn_supported(d)) return 0;
  return Scratch(d, nullptr, want_weights).bytes;
}
int64_t bf16_saved_offset(const Dims& d, const char* name) { return f32_saved_offset(d, name); }

int bf16_forward(const Dims& d, const float* x, const aaconv_params* p, float* y, float* weights, void* saved,
                 void* scratch, cudaStream_t st) {
  AACONV_TRY(tc_attn_supported(d));
  AACONV_TRY(aug_supported(d));
  Scratch w(d, scratch, weights != nullptr);
  float* q = at<float>(saved, f32_saved_offset(d, "q"));
  float* k = at<float>(saved, f32_saved_offset(d, "k"));
  float* v = at<float>(saved, f32_saved_offset(d, "v"));
  float* o = at<float>(saved, f32_saved_offset(d, "o"));
  float* lse = at<float>(saved, f32_saved_offset(d, "lse"));
  if (w.gemm_ok) {
    AACONV_TRY(tc_fprop(d, w.gemm, x, p->conv_w, p->qkv_w, y, q, k, v, st));
  } else {   // geometry outside the TMA tiling (e.g. rows wider than 128 pixels): FFMA implicit GEMM
    AACONV_TRY(f32_conv_fwd(d, x, p->conv_w, y, st));
    AACONV_TRY(f32_qkv_fwd(d, x, p->qkv_w, q, k, v, st));
  }
  AACONV_TRY(tc_attn_fwd(d, q, k, v, p->key_rel_w, p->key_rel_h, w.fwd_operands, o, lse, st));
  if (weights) {   // visualise path only: exact fp32 map (own fp32 statistics), independent of the bf16 kernel
    AACONV_TRY(f32_rel_fwd(d, q, p->key_rel_w, p->key_rel_h, w.rw, w.rh, st));
    AACONV_TRY(f32_attn_fwd(d, q, k, v, w.rw, w.rh, w.o_tmp, w.lse_tmp, st));
    AACONV_TRY(f32_attn_weights(d, q, k, w.rw, w.rh, w.lse_tmp, weights, st));
  }
  AACONV_TRY(f32_out_fwd(d, o, p->out_w, y, st));
  return 0;
}

int bf16_backward(const Dims& d, const float* x, const aaconv_params* p, const float* dy, const void* saved,
                  void* scratch, float* dx, const aaconv_param_grads* g, cudaStream_t st) {
  AACONV_TRY(aug_supported(d));
  Scratch w(d, scratch, 0);
  const AugLayout a = aug_layout(d);
  const float* q = at<float>(saved, f32_saved_offset(d, "q"));
  const float* k = at<float>(saved, f32_saved_offset(d, "k"));
  const float* v = at<float>(saved, f32_saved_offset(d, "v"));
  const float* o = at<float>(saved, f32_saved_offset(d, "o"));
  const float* lse = at<float>(saved, f32_saved_offset(d, "lse"));
  AACONV_TRY(f32_out_bwd(d, dy, o, p->out_w, w.d_o, g->out_w, w.partial, st));
  AACONV_TRY(f32_delta(d, w.d_o, o, w.delta, st));
  AACONV_TRY(aug_build(d, 1, q, k, v, p->key_rel_w, p->key_rel_h, lse, w.d_o, w.delta, w.qa, w.ka, st));
  AACONV_TRY(tc_attn_bwd(d, w.qa, w.ka, w.dqa, w.dk, w.dv, st));
  if (d.relative) {
    if (g->key_rel_w) AACONV_TRY(aug_rel_weight_grad(d, q, w.dqa, a.KD, 0, g->key_rel_w, w.partial, st));
    if (g->key_rel_h) AACONV_TRY(aug_rel_weight_grad(d, q, w.dqa, a.KD, 1, g->key_rel_h, w.partial, st));
  }
  AACONV_TRY(aug_bwd_dq(d, w.dqa, p->key_rel_w, p->key_rel_h, w.dq, st));
  if (w.gemm_ok) {
    AACONV_TRY(tc_pack_grads(d, w.gemm, dy, w.dq, w.dk, w.dv, st));
    if (dx) AACONV_TRY(tc_dgrad(d, w.gemm, p->conv_w, p->qkv_w, dx, st));
    if (g->conv_w || g->qkv_w) {
      if (tc_wgrad_supported(d) == 0) {
        AACONV_TRY(pack_nhwc_bf16(x, w.gemm.xh, d.B, d.Cin, w.gemm.CinK, d.Hin * d.Win, st));
        AACONV_TRY(tc_wgrad(d, w.gemm, d.Cc ? g->conv_w : nullptr, g->qkv_w, w.partial, st));
      } else {
        if (d.Cc) AACONV_TRY(f32_conv_bwd(d, x, p->conv_w, dy, nullptr, g->conv_w, w.partial, st));
        AACONV_TRY(f32_qkv_bwd(d, x, p->qkv_w, w.dq, w.dk, w.dv, g->qkv_w, nullptr, 0, w.partial, st));
      }
    }
    return 0;
  }
  if (d.Cc) {
    AACONV_TRY(f32_conv_bwd(d, x, p->conv_w, dy, dx, g->conv_w, w.partial, st));
  } else if (dx) {
    AACONV_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)d.B * d.Cin * d.Hin * d.Win, st));
  }
  AACONV_TRY(f32_qkv_bwd(d, x, p->qkv_w, w.dq, w.dk, w.dv, g->qkv_w, dx, /*accumulate=*/1, w.partial, st));
  return 0;
}

}  // namespace aaconv
