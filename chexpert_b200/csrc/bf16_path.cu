// Orchestration of the bf16 tensor-core path.  Stages not yet moved to tcgen05 run the fp32 FFMA kernels
// (same layouts), so the path is always complete; DESIGN.md lists which stage runs where.
#include <algorithm>
#include <cstring>
#include "bf16_path.cuh"
#include "fp32_path.cuh"

namespace aaconv {

namespace {
struct Scratch {
  float *d_o, *delta, *dq, *dk, *dv, *partial, *relpart, *dqa;
  float* xn;        // fp32 NCHW relu((x-mean)*rstd) / fp32 copy of a bf16 x: only when an FFMA fallback needs x itself
  void* dxraw;      // gradient wrt the AAConv2d input before the fused InstanceNorm + ReLU adjoint / type conversion
  int dxraw_bf16;
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
    partial = c.take<float>(std::max(std::max(f32_partial_floats(d), tc_gemm_supported(d) == 0 ? tc_wgrad_partial_floats(d) : 0),
                                     out_bwd_patch_partial_floats(d)));
    relpart = c.take<float>(rel_bwd_partial_floats(d));
    dqa = c.take<float>(rows * a.KD);
    gemm_ok = tc_gemm_supported(d) == 0;
    {
      char* gb = c.take<char>(gemm_ok ? tc_gemm_bufs(d, nullptr).bytes : 0);
      gemm = tc_gemm_bufs(d, gb);
    }
    (void)want_weights;   // the attention map is computed from the saved operands: no extra scratch
    const size_t nx = (size_t)d.B * d.Cin * d.Hin * d.Win;
    const bool typed = d.x_bf16 || d.fuse_in;
    const bool ffma_x = !gemm_ok || tc_wgrad_supported(d) != 0;
    xn = (typed && ffma_x) ? c.take<float>(nx) : nullptr;
    // the tcgen05 dgrad writes the raw gradient in the type of x (bf16 under autocast); the FFMA fallback in fp32
    dxraw_bf16 = (gemm_ok && d.x_bf16) ? 1 : 0;
    dxraw = typed ? c.take<char>(nx * (dxraw_bf16 ? 2 : 4)) : nullptr;
    bytes = c.off;
  }
};
template <class T>
T* at(const void* base, int64_t off) { return reinterpret_cast<T*>(static_cast<char*>(const_cast<void*>(base)) + off); }

// saved block = the fp32 block (q,k,v,o,lse; same offsets as the fp32 path) followed by the augmented operands
struct SavedAug {
  void *qa, *ka, *xh;       // xh: channels-last bf16 copy of x (fprop operand, reused by wgrad)
  float* stats;             // (B*Cin) x (mean, rstd) of the fused InstanceNorm prologue
  size_t bytes;
  SavedAug(const Dims& d, const void* base) {
    Carver c(const_cast<void*>(base));
    c.take<char>(f32_saved_bytes(d));
    const size_t rows = (size_t)d.BN * d.L;
    const AugLayout a = aug_layout(d);
    qa = c.take<uint16_t>(rows * a.KP);
    ka = c.take<uint16_t>(rows * a.KP);
    xh = c.take<uint16_t>(tc_gemm_supported(d) == 0 ? (size_t)d.B * d.Hin * d.Win * (cdiv(d.Cin, 64) * 64) : 0);
    stats = c.take<float>(d.fuse_in ? (size_t)d.B * d.Cin * 2 : 0);
    bytes = c.off;
  }
};
}  // namespace

size_t bf16_saved_bytes(const Dims& d) { return SavedAug(d, nullptr).bytes; }
size_t bf16_scratch_bytes(const Dims& d, int want_weights) {
  if (aug_supported(d) || tc_attn_supported(d)) return 0;
  return Scratch(d, nullptr, want_weights).bytes;
}
int64_t bf16_saved_offset(const Dims& d, const char* name) { return f32_saved_offset(d, name); }

int bf16_forward(const Dims& d, const void* xin, const aaconv_params* p, void* y, float* weights, void* saved,
                 void* scratch, cudaStream_t st) {
  AACONV_TRY(tc_attn_supported(d));
  AACONV_TRY(aug_supported(d));
  Scratch w(d, scratch, weights != nullptr);
  float* q = at<float>(saved, f32_saved_offset(d, "q"));
  float* k = at<float>(saved, f32_saved_offset(d, "k"));
  float* v = at<float>(saved, f32_saved_offset(d, "v"));
  float* o = at<float>(saved, f32_saved_offset(d, "o"));
  float* lse = at<float>(saved, f32_saved_offset(d, "lse"));
  SavedAug sa(d, saved);
  NvtxRange r_fwd("aaconv.forward");
  if (d.fuse_in) { NvtxRange r("aaconv.prologue"); AACONV_TRY(in_stats(d, xin, sa.stats, st)); }   // InstanceNorm statistics (attn_aug_conv.py:438)
  const float* x = static_cast<const float*>(xin);                // fp32 view of the (normalised) input for the FFMA fallbacks
  if (w.xn && !w.gemm_ok) { AACONV_TRY(in_relu_apply(d, xin, sa.stats, w.xn, st)); x = w.xn; }
  if (w.gemm_ok) {
    NvtxRange r("aaconv.fprop");
    w.gemm.xh = sa.xh;
    // layout + precision pack of the GEMM operand, with relu((x-mean)*rstd) applied on the way (attn_aug_conv.py:438-439)
    AACONV_TRY(pack_x(xin, d.x_bf16, d.fuse_in ? sa.stats : nullptr, sa.xh, d.B, d.Cin, w.gemm.CinK, d.Hin * d.Win, st));
    AACONV_TRY(tc_fprop(d, w.gemm, p->conv_w, p->qkv_w, y, q, k, v, st));
  } else {   // geometry outside the TMA tiling (e.g. rows wider than 128 pixels): FFMA implicit GEMM
    AACONV_TRY(f32_conv_fwd(d, x, p->conv_w, y, st));
    AACONV_TRY(f32_qkv_fwd(d, x, p->qkv_w, q, k, v, st));
  }
  { NvtxRange r("aaconv.aug_build"); AACONV_TRY(aug_build_fwd(d, q, k, v, p->key_rel_w, p->key_rel_h, sa.qa, sa.ka, st)); }
  {
    NvtxRange r("aaconv.attention");
    if (cc_attn_supported(d) == 0) AACONV_TRY(cc_attn_fwd(d, sa.qa, sa.ka, v, o, lse, st));
    else AACONV_TRY(tc_attn_fwd(d, sa.qa, sa.ka, o, lse, st));
  }
  NvtxRange r_epi("aaconv.epilogue");
  // visualise path only: the bf16 kernels' own probabilities -- their bf16 operands, their lse (attn_aug_conv.py:87)
  if (weights) AACONV_TRY(aug_weights(d, sa.qa, sa.ka, lse, weights, st));
  AACONV_TRY(out_proj_fwd(d, o, p->out_w, y, st));
  return 0;
}

int bf16_backward(const Dims& d, const void* xin, const aaconv_params* p, const float* dy, void* saved,
                  void* scratch, void* dx_out, const aaconv_param_grads* g, cudaStream_t st) {
  AACONV_TRY(aug_supported(d));
  Scratch w(d, scratch, 0);
  // typed boundary / fused prologue: the kernels below produce the gradient wrt the AAConv2d input in w.dxraw; the last step
  // turns it into dx (InstanceNorm + ReLU adjoint, or a plain type conversion).  Otherwise they write dx directly.
  // (a bf16 x without the fused prologue whose raw gradient is already bf16 needs no second pass: dgrad writes dx itself)
  const bool typed = d.fuse_in || (d.x_bf16 && !w.dxraw_bf16);
  void* const dxv = !dx_out ? nullptr : (typed ? w.dxraw : dx_out);
  float* const dx = static_cast<float*>(dxv);                      // fp32 view for the FFMA fallbacks
  const int dx_bf16 = typed ? w.dxraw_bf16 : d.x_bf16;
  const float* x = static_cast<const float*>(xin);
  const AugLayout a = aug_layout(d);
  const float* q = at<float>(saved, f32_saved_offset(d, "q"));
  const float* k = at<float>(saved, f32_saved_offset(d, "k"));
  const float* v = at<float>(saved, f32_saved_offset(d, "v"));
  const float* o = at<float>(saved, f32_saved_offset(d, "o"));
  const float* lse = at<float>(saved, f32_saved_offset(d, "lse"));
  SavedAug sa(d, saved);
  NvtxRange r_bwd("aaconv.backward");
  if (w.xn) { AACONV_TRY(in_relu_apply(d, xin, sa.stats, w.xn, st)); x = w.xn; }
  // the backward-only columns of Qa (-lse, dO, -delta) are filled in place; idempotent, so a retained graph may
  // run backward again
  if (out_bwd_patch_supported(d) == 0) {     // out_proj adjoint and the patch in one pass over the pixels
    AACONV_TRY(out_bwd_patch(d, dy, o, lse, p->out_w, w.d_o, w.delta, sa.qa, g->out_w, w.partial, st));
  } else {                                   // wider values (Transition 2 / 3): data part + patch fused, dWout as a split-K GEMM
    AACONV_TRY(out_bwd_data_patch(d, dy, o, lse, p->out_w, w.d_o, w.delta, sa.qa, st));
    AACONV_TRY(f32_out_bwd_weight(d, dy, o, g->out_w, w.partial, st));
  }
  // fast path: the attention-backward kernels write dq*scale, dk, dv as bf16 straight into the packed (B*L, KPq)
  // operand of the projection dgrad/wgrad GEMMs; its padding columns must be finite (they meet zero weights)
  const bool direct = w.gemm_ok && rel_bwd_supported(d) == 0 && tc_wgrad_supported(d) == 0;
  if (direct) AACONV_CUDA_OK(cudaMemsetAsync(w.gemm.dqkvh, 0, sizeof(uint16_t) * (size_t)d.B * d.L * w.gemm.KPq, st));
  {
    NvtxRange r("aaconv.attention_bwd");
    if (cc_attn_supported(d) == 0)
      AACONV_TRY(cc_attn_bwd(d, sa.qa, sa.ka, v, w.d_o, w.delta, w.dqa, w.dk, w.dv, direct ? w.gemm.dqkvh : nullptr, w.gemm.KPq, st));
    else
      AACONV_TRY(tc_attn_bwd(d, sa.qa, sa.ka, w.dqa, w.dk, w.dv, direct ? w.gemm.dqkvh : nullptr, w.gemm.KPq, st));
  }
  NvtxRange r_rest("aaconv.rel_bwd+dgrad+wgrad");
  if (direct) {
    AACONV_TRY(rel_bwd(d, w.dqa, q, p->key_rel_w, p->key_rel_h, nullptr, w.gemm.dqkvh, w.gemm.KPq, g->key_rel_w, g->key_rel_h,
                       w.relpart, st));
  } else {
    if (d.relative) {
      if (g->key_rel_w) AACONV_TRY(aug_rel_weight_grad(d, q, w.dqa, a.KD, 0, g->key_rel_w, w.partial, st));
      if (g->key_rel_h) AACONV_TRY(aug_rel_weight_grad(d, q, w.dqa, a.KD, 1, g->key_rel_h, w.partial, st));
    }
    AACONV_TRY(aug_bwd_dq(d, w.dqa, p->key_rel_w, p->key_rel_h, w.dq, st));
  }
  if (w.gemm_ok) {
    AACONV_TRY(tc_pack_grads(d, w.gemm, dy, direct ? nullptr : w.dq, w.dk, w.dv, st));
    if (dxv) AACONV_TRY(tc_dgrad(d, w.gemm, p->conv_w, p->qkv_w, dxv, dx_bf16, st));
    if (g->conv_w || g->qkv_w) {
      if (tc_wgrad_supported(d) == 0) {
        w.gemm.xh = sa.xh;             // packed by forward
        AACONV_TRY(tc_wgrad(d, w.gemm, d.Cc ? g->conv_w : nullptr, g->qkv_w, w.partial, st));
      } else {
        if (d.Cc) AACONV_TRY(f32_conv_bwd(d, x, p->conv_w, dy, nullptr, g->conv_w, w.partial, st));
        AACONV_TRY(f32_qkv_bwd(d, x, p->qkv_w, w.dq, w.dk, w.dv, g->qkv_w, nullptr, 0, w.partial, st));
      }
    }
    if (typed && dx_out) AACONV_TRY(in_relu_bwd(d, xin, w.dxraw, w.dxraw_bf16, sa.stats, dx_out, st));
    return 0;
  }
  if (d.Cc) {
    AACONV_TRY(f32_conv_bwd(d, x, p->conv_w, dy, dx, g->conv_w, w.partial, st));
  } else if (dx) {
    AACONV_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)d.B * d.Cin * d.Hin * d.Win, st));
  }
  AACONV_TRY(f32_qkv_bwd(d, x, p->qkv_w, w.dq, w.dk, w.dv, g->qkv_w, dx, /*accumulate=*/1, w.partial, st));
  if (typed && dx_out) AACONV_TRY(in_relu_bwd(d, xin, w.dxraw, w.dxraw_bf16, sa.stats, dx_out, st));
  return 0;
}

}  // namespace aaconv
