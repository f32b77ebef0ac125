// Orchestration of the fp32 path: AAConv2d.forward (attn_aug_conv.py:65-97) and its adjoint as a fixed
// sequence of launches on one stream.  All buffers are caller-owned (see include/aaconv_b200.h).
#include <cstring>
#include "fp32_path.cuh"

namespace aaconv {

namespace {
struct Saved {
  float *q, *k, *v, *o, *lse;
  size_t bytes;
  Saved(const Dims& d, void* base) {
    Carver c(base);
    const size_t rows = (size_t)d.BN * d.L;
    q = c.take<float>(rows * d.dkh);
    k = c.take<float>(rows * d.dkh);
    v = c.take<float>(rows * d.dvh);
    o = c.take<float>(rows * d.dvh);
    lse = c.take<float>(rows);
    bytes = c.off;
  }
};
struct Scratch {
  float *rw, *rh, *drw, *drh, *d_o, *delta, *dq, *dk, *dv, *partial;
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
    bytes = c.off;
  }
};
}  // namespace

size_t f32_saved_bytes(const Dims& d) { return Saved(d, nullptr).bytes; }
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

int f32_forward(const Dims& d, const float* x, const aaconv_params* p, float* y, float* weights, void* saved,
                void* scratch, cudaStream_t st) {
  Saved s(d, saved);
  Scratch w(d, scratch);
  AACONV_TRY(f32_conv_fwd(d, x, p->conv_w, y, st));
  AACONV_TRY(f32_qkv_fwd(d, x, p->qkv_w, s.q, s.k, s.v, st));
  AACONV_TRY(f32_rel_fwd(d, s.q, p->key_rel_w, p->key_rel_h, w.rw, w.rh, st));
  AACONV_TRY(f32_attn_fwd(d, s.q, s.k, s.v, w.rw, w.rh, s.o, s.lse, st));
  if (weights) AACONV_TRY(f32_attn_weights(d, s.q, s.k, w.rw, w.rh, s.lse, weights, st));
  AACONV_TRY(f32_out_fwd(d, s.o, p->out_w, y, st));
  return 0;
}

int f32_backward(const Dims& d, const float* x, const aaconv_params* p, const float* dy, void* saved,
                 void* scratch, float* dx, const aaconv_param_grads* g, cudaStream_t st) {
  Saved s(d, const_cast<void*>(saved));
  Scratch w(d, scratch);
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
  return 0;
}

}  // namespace aaconv
