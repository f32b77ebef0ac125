// Orchestration of the fp32 path: AAConv2d.forward (attn_aug_conv.py:65-97) and its adjoint as a fixed
// sequence of launches on one stream.  All buffers are caller-owned (see include/aaconv_b200.h).
#include <cstring>
#include "fp32_path.cuh"
#include "bf16_path.cuh"   // prologue.cu: in_stats / in_relu_apply / in_relu_bwd

namespace aaconv {

namespace {
struct Saved {
  float *q, *k, *v, *o, *lse, *stats;
  size_t bytes;
  Saved(const Dims& d, void* base) {
    Carver c(base);
    const size_t rows = (size_t)d.BN * d.L;
    q = c.take<float>(rows * d.dkh);
    k = c.take<float>(rows * d.dkh);
    v = c.take<float>(rows * d.dvh);
    o = c.take<float>(rows * d.dvh);
    lse = c.take<float>(rows);
    bytes = c.off;                       // the bf16 path appends its own blocks after these (same offsets for q..lse)
    stats = nullptr;
  }
};
// (B*Cin) x (mean, rstd) of the fused InstanceNorm prologue live at the END of the fp32 saved block
struct SavedStats {
  float* stats;
  size_t bytes;
  SavedStats(const Dims& d, void* base, size_t off) {
    Carver c(base);
    c.off = off;
    stats = c.take<float>(d.fuse_in ? (size_t)d.B * d.Cin * 2 : 0);
    bytes = c.off;
  }
};
struct Scratch {
  float *rw, *rh, *drw, *drh, *d_o, *delta, *dq, *dk, *dv, *partial, *xn, *dxraw;
  size_t bytes;
  Scratch(const Dims& d, void* base) {
    Carver c(base);
    const size_t rows = (size_t)d.BN * d.L;
    rw = c.take<float>(d.relative ? rows * d.RW : 0);
    rh = c.take<float>(d.relative ? rows * d.RH : 0);
    drw = c.take<float>(d.relative ? rows * d.RW : 0);
    drh = c.take<float>(d.relative ? rows * d.RH : 0);
    d_o = c.take<float>(rows * d.dvh);
    delta = c.take<float>(rows);
    dq = c.take<float>(rows * d.dkh);
    dk = c.take<float>(rows * d.dkh);
    dv = c.take<float>(rows * d.dvh);
    partial = c.take<float>(f32_partial_floats(d));
    const size_t nx = (size_t)d.B * d.Cin * d.Hin * d.Win;
    const bool typed = d.x_bf16 || d.fuse_in;
    xn = typed ? c.take<float>(nx) : nullptr;        // fp32 relu((x-mean)*rstd) / fp32 copy of a bf16 x
    dxraw = typed ? c.take<float>(nx) : nullptr;     // gradient wrt the AAConv2d input before the prologue's adjoint
    bytes = c.off;
  }
};
}  // namespace

size_t f32_saved_bytes(const Dims& d) { return Saved(d, nullptr).bytes; }
size_t f32_saved_bytes_io(const Dims& d) { return SavedStats(d, nullptr, Saved(d, nullptr).bytes).bytes; }
size_t f32_scratch_bytes(const Dims& d) { return Scratch(d, nullptr).bytes; }

int64_t f32_saved_offset(const Dims& d, const char* name) {
  Saved s(d, reinterpret_cast<void*>(uintptr_t(256)));   // non-null dummy base -> pointer differences
  const char* base = reinterpret_cast<const char*>(uintptr_t(256));
  auto off = [&](const float* p) { return (int64_t)(reinterpret_cast<const char*>(p) - base); };
  if (!strcmp(name, "q")) return off(s.q);
  if (!strcmp(name, "k")) return off(s.k);
  if (!strcmp(name, "v")) return off(s.v);
  if (!strcmp(name, "o")) return off(s.o);
  if (!strcmp(name, "lse")) return off(s.lse);
  return -1;
}

int f32_forward(const Dims& d, const void* xin, const aaconv_params* p, void* y, float* weights, void* saved,
                void* scratch, cudaStream_t st) {
  Saved s(d, saved);
  SavedStats ss(d, saved, s.bytes);
  Scratch w(d, scratch);
  const float* x = static_cast<const float*>(xin);
  if (d.fuse_in) AACONV_TRY(in_stats(d, xin, ss.stats, st));      // InstanceNorm2d + ReLU of the Transition (attn_aug_conv.py:438-439)
  if (w.xn) { AACONV_TRY(in_relu_apply(d, xin, ss.stats, w.xn, st)); x = w.xn; }
  AACONV_TRY(f32_conv_fwd(d, x, p->conv_w, y, st));
  AACONV_TRY(f32_qkv_fwd(d, x, p->qkv_w, s.q, s.k, s.v, st));
  AACONV_TRY(f32_rel_fwd(d, s.q, p->key_rel_w, p->key_rel_h, w.rw, w.rh, st));
  AACONV_TRY(f32_attn_fwd(d, s.q, s.k, s.v, w.rw, w.rh, s.o, s.lse, st));
  if (weights) AACONV_TRY(f32_attn_weights(d, s.q, s.k, w.rw, w.rh, s.lse, weights, st));
  AACONV_TRY(f32_out_fwd(d, s.o, p->out_w, y, st));
  return 0;
}

int f32_backward(const Dims& d, const void* xin, const aaconv_params* p, const float* dy, void* saved,
                 void* scratch, void* dx_out, const aaconv_param_grads* g, cudaStream_t st) {
  Saved s(d, const_cast<void*>(saved));
  SavedStats ss(d, saved, s.bytes);
  Scratch w(d, scratch);
  const float* x = static_cast<const float*>(xin);
  if (w.xn) { AACONV_TRY(in_relu_apply(d, xin, ss.stats, w.xn, st)); x = w.xn; }
  const bool typed = d.x_bf16 || d.fuse_in;
  float* const dx = !dx_out ? nullptr : (typed ? w.dxraw : static_cast<float*>(dx_out));
  AACONV_TRY(f32_out_bwd(d, dy, s.o, p->out_w, w.d_o, g->out_w, w.partial, st));
  AACONV_TRY(f32_delta(d, w.d_o, s.o, w.delta, st));
  AACONV_TRY(f32_rel_fwd(d, s.q, p->key_rel_w, p->key_rel_h, w.rw, w.rh, st));
  AACONV_TRY(f32_attn_bwd(d, s.q, s.k, s.v, w.rw, w.rh, s.lse, w.d_o, w.delta, w.dq, w.dk, w.dv, w.drw, w.drh, st));
  if (d.relative) {
    if (g->key_rel_w) AACONV_TRY(f32_rel_weight_grad(d, s.q, w.drw, d.RW, g->key_rel_w, w.partial, st));
    if (g->key_rel_h) AACONV_TRY(f32_rel_weight_grad(d, s.q, w.drh, d.RH, g->key_rel_h, w.partial, st));
    AACONV_TRY(f32_rel_bwd_dq(d, p->key_rel_w, p->key_rel_h, w.drw, w.drh, w.dq, st));
  }
  if (d.Cc) {
    AACONV_TRY(f32_conv_bwd(d, x, p->conv_w, dy, dx, g->conv_w, w.partial, st));
  } else if (dx) {
    AACONV_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)d.B * d.Cin * d.Hin * d.Win, st));
  }
  AACONV_TRY(f32_qkv_bwd(d, x, p->qkv_w, w.dq, w.dk, w.dv, g->qkv_w, dx, /*accumulate=*/1, w.partial, st));
  if (typed && dx_out) AACONV_TRY(in_relu_bwd(d, xin, w.dxraw, 0, ss.stats, dx_out, st));
  return 0;
}

}  // namespace aaconv
