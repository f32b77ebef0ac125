// fp32 (FFMA) dense contractions of the AAConv2d path, expressed as index maps over simt_gemm.cuh.
// Reference rows: a2 qkv projection, a9 out_proj, a10 conv branch, and their adjoints.
#include "fp32_path.cuh"
#include "simt_gemm.cuh"
#include "bf16_path.cuh"

namespace aaconv {

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                     int count, int splits, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partial[(size_t)k * count + i];
  out[i] = accumulate ? out[i] + s : s;
}

static int splitk_reduce(const float* partial, float* out, int count, int splits, cudaStream_t st) {
  splitk_reduce_kernel<<<cdiv(count, 256), 256, 0, AACONV_ST(st)>>>(partial, out, count, splits, 0);
  AACONV_LAUNCH_OK("splitk_reduce");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// conv branch fprop: y[:, coff:coff+Cc] = conv2d(x, w)           (attn_aug_conv.py:34,95)
// ------------------------------------------------------------------------------------------------
struct ConvFwdP {
  int M, N, K, k_chunk;
  const float* x; const float* w; void* y;
  int Cin, Hin, Win, Ho, Wo, ks, stride, pad, dil, coff, y_bf16; long long y_bs;
  static constexpr bool A_K_CONTIG = false, B_K_CONTIG = true;
  struct ARow { const float* xb; int h0, w0; };
  struct BCol { const float* wr; };
  __device__ ARow a_row(int m) const {
    const int j = m % Wo, t = m / Wo, i = t % Ho, b = t / Ho;
    return {x + (size_t)b * Cin * Hin * Win, i * stride - pad, j * stride - pad};
  }
  __device__ float a(const ARow& r, int k) const {
    const int kk = ks * ks, c = k / kk, rem = k - c * kk, kh = rem / ks, kw = rem - kh * ks;
    const int h = r.h0 + kh * dil, w_ = r.w0 + kw * dil;
    return (h >= 0 && h < Hin && w_ >= 0 && w_ < Win) ? __ldg(r.xb + ((size_t)c * Hin + h) * Win + w_) : 0.f;
  }
  __device__ BCol b_col(int n) const { return {w + (size_t)n * K}; }
  __device__ float b(const BCol& c, int k) const { return __ldg(c.wr + k); }
  __device__ void store(int m, int n, float v, int) const {
    const int j = m % Wo, t = m / Wo, i = t % Ho, b = t / Ho;
    store_y(y, (size_t)b * y_bs + ((size_t)(coff + n) * Ho + i) * Wo + j, v, y_bf16);
  }
};

int f32_conv_fwd(const Dims& d, const float* x, const float* w, void* y, cudaStream_t st) {
  if (d.Cc == 0) return 0;
  ConvFwdP p;
  p.M = d.B * d.L; p.N = d.Cc; p.K = d.Cin * d.ks * d.ks; p.k_chunk = p.K;
  p.x = x; p.w = w; p.y = y;
  p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.Ho = d.H; p.Wo = d.W;
  p.ks = d.ks; p.stride = d.stride; p.pad = d.pad; p.dil = d.dil; p.coff = 0; p.y_bs = d.y_bs; p.y_bf16 = d.y_bf16;
  return launch_simt_gemm(p, 1, st, "conv_fwd_f32");
}

// ------------------------------------------------------------------------------------------------
// qkv projection + head split + q scale                          (attn_aug_conv.py:67-73)
// ------------------------------------------------------------------------------------------------
struct QkvFwdP {
  int M, N, K, k_chunk;
  const float* x; const float* w; float* q; float* kk_; float* v;
  int Cin, Hin, Win, Wo, L, stride, dk, dkh, dvh, nh;
  float qscale;
  static constexpr bool A_K_CONTIG = false, B_K_CONTIG = true;
  struct ARow { const float* px; };
  struct BCol { const float* wr; };
  __device__ ARow a_row(int m) const {
    const int b = m / L, l = m - b * L, i = l / Wo, j = l - i * Wo;
    return {x + (size_t)b * Cin * Hin * Win + (size_t)(i * stride) * Win + j * stride};
  }
  __device__ float a(const ARow& r, int k) const { return __ldg(r.px + (size_t)k * Hin * Win); }
  __device__ BCol b_col(int n) const { return {w + (size_t)n * K}; }
  __device__ float b(const BCol& c, int k) const { return __ldg(c.wr + k); }
  __device__ void store(int m, int n, float val, int) const {
    const int b = m / L, l = m - b * L;
    if (n < dk) {
      const int h = n / dkh, e = n - h * dkh;
      q[((size_t)(b * nh + h) * L + l) * dkh + e] = val * qscale;
    } else if (n < 2 * dk) {
      const int c = n - dk, h = c / dkh, e = c - h * dkh;
      kk_[((size_t)(b * nh + h) * L + l) * dkh + e] = val;
    } else {
      const int c = n - 2 * dk, h = c / dvh, e = c - h * dvh;
      v[((size_t)(b * nh + h) * L + l) * dvh + e] = val;
    }
  }
};

int f32_qkv_fwd(const Dims& d, const float* x, const float* w, float* q, float* k, float* v, cudaStream_t st) {
  QkvFwdP p;
  p.M = d.B * d.L; p.N = d.Nqkv; p.K = d.Cin; p.k_chunk = p.K;
  p.x = x; p.w = w; p.q = q; p.kk_ = k; p.v = v;
  p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.Wo = d.W; p.L = d.L; p.stride = d.stride;
  p.dk = d.dk; p.dkh = d.dkh; p.dvh = d.dvh; p.nh = d.nh; p.qscale = d.qscale;
  return launch_simt_gemm(p, 1, st, "qkv_fwd_f32");
}

// ------------------------------------------------------------------------------------------------
// head combine + out_proj + concat write                         (attn_aug_conv.py:89-95)
// ------------------------------------------------------------------------------------------------
struct OutFwdP {
  int M, N, K, k_chunk;
  const float* o; const float* w; void* y;
  int L, nh, dvh, coff, y_bf16; long long y_bs;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
  struct ARow { int b, l; };
  struct BCol { const float* wr; };
  __device__ ARow a_row(int m) const { const int b = m / L; return {b, m - b * L}; }
  __device__ float a(const ARow& r, int k) const {
    const int h = k / dvh, e = k - h * dvh;
    return __ldg(o + ((size_t)(r.b * nh + h) * L + r.l) * dvh + e);
  }
  __device__ BCol b_col(int n) const { return {w + (size_t)n * K}; }
  __device__ float b(const BCol& c, int k) const { return __ldg(c.wr + k); }
  __device__ void store(int m, int n, float v, int) const {
    const int b = m / L, l = m - b * L;
    store_y(y, (size_t)b * y_bs + (size_t)(coff + n) * L + l, v, y_bf16);
  }
};

int f32_out_fwd(const Dims& d, const float* o, const float* w, void* y, cudaStream_t st) {
  OutFwdP p;
  p.M = d.B * d.L; p.N = d.dv; p.K = d.dv; p.k_chunk = p.K;
  p.o = o; p.w = w; p.y = y; p.L = d.L; p.nh = d.nh; p.dvh = d.dvh; p.coff = d.Cc; p.y_bs = d.y_bs; p.y_bf16 = d.y_bf16;
  return launch_simt_gemm(p, 1, st, "out_fwd_f32");
}

// dO[b,n,l,e] = sum_c Wout[c, n*dvh+e] * dy[b, coff+c, l]
struct OutBwdDataP {
  int M, N, K, k_chunk;
  const float* dy; const float* w; float* d_o;
  int L, nh, dvh, Ctot, coff;
  static constexpr bool A_K_CONTIG = false, B_K_CONTIG = false;
  struct ARow { const float* p; };
  struct BCol { int n; };
  __device__ ARow a_row(int m) const {
    const int b = m / L, l = m - b * L;
    return {dy + ((size_t)b * Ctot + coff) * L + l};
  }
  __device__ float a(const ARow& r, int k) const { return __ldg(r.p + (size_t)k * L); }
  __device__ BCol b_col(int n) const { return {n}; }
  __device__ float b(const BCol& c, int k) const { return __ldg(w + (size_t)k * N + c.n); }
  __device__ void store(int m, int n, float v, int) const {
    const int b = m / L, l = m - b * L, h = n / dvh, e = n - h * dvh;
    d_o[((size_t)(b * nh + h) * L + l) * dvh + e] = v;
  }
};

// dWout[c, c'] = sum_{b,l} dy[b, coff+c, l] * o[b, n(c'), l, e(c')]
struct OutBwdWeightP {
  int M, N, K, k_chunk;
  const float* dy; const float* o; float* partial;
  int L, nh, dvh, Ctot, coff;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
  struct ARow { int c; };
  struct BCol { int h, e; };
  __device__ ARow a_row(int m) const { return {m}; }
  __device__ float a(const ARow& r, int k) const {
    const int b = k / L, l = k - b * L;
    return __ldg(dy + ((size_t)b * Ctot + coff + r.c) * L + l);
  }
  __device__ BCol b_col(int n) const { const int h = n / dvh; return {h, n - h * dvh}; }
  __device__ float b(const BCol& c, int k) const {
    const int b = k / L, l = k - b * L;
    return __ldg(o + ((size_t)(b * nh + c.h) * L + l) * dvh + c.e);
  }
  __device__ void store(int m, int n, float v, int split) const {
    partial[(size_t)split * M * N + (size_t)m * N + n] = v;
  }
};

int f32_out_bwd(const Dims& d, const float* dy, const float* o, const float* w, float* d_o, float* dw,
                float* partial, cudaStream_t st) {
  OutBwdDataP p;
  p.M = d.B * d.L; p.N = d.dv; p.K = d.dv; p.k_chunk = p.K;
  p.dy = dy; p.w = w; p.d_o = d_o; p.L = d.L; p.nh = d.nh; p.dvh = d.dvh; p.Ctot = d.Cout; p.coff = d.Cc;
  AACONV_TRY(launch_simt_gemm(p, 1, st, "out_bwd_data_f32"));
  return f32_out_bwd_weight(d, dy, o, dw, partial, st);
}

int f32_out_bwd_weight(const Dims& d, const float* dy, const float* o, float* dw, float* partial, cudaStream_t st) {
  if (dw) {
    OutBwdWeightP q;
    q.M = d.dv; q.N = d.dv; q.K = d.B * d.L;
    int splits = pick_splits(q.M, q.N, q.K);
    q.k_chunk = chunk_for(q.K, splits);
    splits = cdiv(q.K, q.k_chunk);
    q.dy = dy; q.o = o; q.partial = partial; q.L = d.L; q.nh = d.nh; q.dvh = d.dvh; q.Ctot = d.Cout; q.coff = d.Cc;
    AACONV_TRY(launch_simt_gemm(q, splits, st, "out_bwd_weight_f32"));
    AACONV_TRY(splitk_reduce(partial, dw, q.M * q.N, splits, st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// qkv projection adjoint.  G(pix, n) = concat(dq*scale, dk, dv) read in place from the head-split grads.
// ------------------------------------------------------------------------------------------------
struct QkvGrad {
  const float* dq; const float* dk; const float* dv;
  int L, nh, dk_, dkh, dvh;
  float qscale;
  struct Col { const float* base; int dh; float scale; };   // base already offset by head and element
  __device__ Col col(int n) const {
    if (n < dk_) { const int h = n / dkh, e = n - h * dkh; return {dq + (size_t)h * L * dkh + e, dkh, qscale}; }
    if (n < 2 * dk_) { const int c = n - dk_, h = c / dkh, e = c - h * dkh; return {dk + (size_t)h * L * dkh + e, dkh, 1.f}; }
    const int c = n - 2 * dk_, h = c / dvh, e = c - h * dvh;
    return {dv + (size_t)h * L * dvh + e, dvh, 1.f};
  }
  __device__ float at(const Col& c, int b, int l) const {
    return __ldg(c.base + ((size_t)b * nh * L + l) * c.dh) * c.scale;
  }
};

// dWqkv[n, cin] = sum_pix G(pix, n) * x[b, cin, i*s, j*s]
struct QkvBwdWeightP {
  int M, N, K, k_chunk;
  QkvGrad g; const float* x; float* partial;
  int Cin, Hin, Win, Wo, L, stride;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
  using ARow = QkvGrad::Col;
  struct BCol { const float* xc; };
  __device__ ARow a_row(int m) const { return g.col(m); }
  __device__ float a(const ARow& r, int k) const { const int b = k / L; return g.at(r, b, k - b * L); }
  __device__ BCol b_col(int n) const { return {x + (size_t)n * Hin * Win}; }
  __device__ float b(const BCol& c, int k) const {
    const int b_ = k / L, l = k - b_ * L, i = l / Wo, j = l - i * Wo;
    return __ldg(c.xc + (size_t)b_ * Cin * Hin * Win + (size_t)(i * stride) * Win + j * stride);
  }
  __device__ void store(int m, int n, float v, int split) const {
    partial[(size_t)split * M * N + (size_t)m * N + n] = v;
  }
};

// dx[b, cin, i*s, j*s] (+)= sum_n G(pix, n) * Wqkv[n, cin]
struct QkvBwdDataP {
  int M, N, K, k_chunk;
  QkvGrad g; const float* w; float* dx;
  int Cin, Hin, Win, Wo, L, stride, accumulate;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = false;
  struct ARow { int b, l; };
  struct BCol { int n; };
  __device__ ARow a_row(int m) const { const int b = m / L; return {b, m - b * L}; }
  __device__ float a(const ARow& r, int k) const { return g.at(g.col(k), r.b, r.l); }
  __device__ BCol b_col(int n) const { return {n}; }
  __device__ float b(const BCol& c, int k) const { return __ldg(w + (size_t)k * N + c.n); }
  __device__ void store(int m, int n, float v, int) const {
    const int b = m / L, l = m - b * L, i = l / Wo, j = l - i * Wo;
    float* p = dx + ((size_t)(b * Cin + n) * Hin + i * stride) * Win + j * stride;
    *p = accumulate ? *p + v : v;
  }
};

int f32_qkv_bwd(const Dims& d, const float* x, const float* w, const float* dq, const float* dk,
                const float* dv, float* dw, float* dx, int dx_accumulate, float* partial, cudaStream_t st) {
  QkvGrad g{dq, dk, dv, d.L, d.nh, d.dk, d.dkh, d.dvh, d.qscale};
  if (dw) {
    QkvBwdWeightP p;
    p.M = d.Nqkv; p.N = d.Cin; p.K = d.B * d.L;
    int splits = pick_splits(p.M, p.N, p.K);
    p.k_chunk = chunk_for(p.K, splits);
    splits = cdiv(p.K, p.k_chunk);
    p.g = g; p.x = x; p.partial = partial;
    p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.Wo = d.W; p.L = d.L; p.stride = d.stride;
    AACONV_TRY(launch_simt_gemm(p, splits, st, "qkv_bwd_weight_f32"));
    AACONV_TRY(splitk_reduce(partial, dw, p.M * p.N, splits, st));
  }
  if (dx) {
    QkvBwdDataP p;
    p.M = d.B * d.L; p.N = d.Cin; p.K = d.Nqkv; p.k_chunk = p.K;
    p.g = g; p.w = w; p.dx = dx;
    p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.Wo = d.W; p.L = d.L; p.stride = d.stride;
    p.accumulate = dx_accumulate;
    AACONV_TRY(launch_simt_gemm(p, 1, st, "qkv_bwd_data_f32"));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// conv branch adjoint
// ------------------------------------------------------------------------------------------------
// dgrad, one launch per input-pixel residue class (h % s, w % s): only the taps that can reach that
// class are enumerated, so no multiply-by-zero work is done for strided convs.
struct ConvBwdDataP {
  int M, N, K, k_chunk;
  const float* dy; const float* w; float* dx;
  int Cin, Hin, Win, Ho, Wo, ks, stride, pad, dil, Ctot, Cc;
  int rh, rw, Hc, Wc, nkh, nkw;
  int khs[8], kws[8];
  static constexpr bool A_K_CONTIG = false, B_K_CONTIG = false;
  struct ARow { const float* dyb; int h, w_; };
  struct BCol { int n; };
  __device__ ARow a_row(int m) const {
    const int wc = m % Wc, t = m / Wc, hc = t % Hc, b = t / Hc;
    return {dy + (size_t)b * Ctot * Ho * Wo, hc * stride + rh, wc * stride + rw};
  }
  __device__ float a(const ARow& r, int k) const {
    const int nt = nkh * nkw, co = k / nt, rem = k - co * nt, a_ = rem / nkw, b_ = rem - a_ * nkw;
    const int i = (r.h + pad - khs[a_] * dil) / stride, j = (r.w_ + pad - kws[b_] * dil) / stride;
    const int ih = r.h + pad - khs[a_] * dil, jw = r.w_ + pad - kws[b_] * dil;
    return (ih >= 0 && jw >= 0 && i < Ho && j < Wo) ? __ldg(r.dyb + ((size_t)co * Ho + i) * Wo + j) : 0.f;
  }
  __device__ BCol b_col(int n) const { return {n}; }
  __device__ float b(const BCol& c, int k) const {
    const int nt = nkh * nkw, co = k / nt, rem = k - co * nt, a_ = rem / nkw, b_ = rem - a_ * nkw;
    return __ldg(w + (((size_t)co * Cin + c.n) * ks + khs[a_]) * ks + kws[b_]);
  }
  __device__ void store(int m, int n, float v, int) const {
    const int wc = m % Wc, t = m / Wc, hc = t % Hc, b = t / Hc;
    dx[((size_t)(b * Cin + n) * Hin + hc * stride + rh) * Win + wc * stride + rw] = v;
  }
};

// wgrad: dW[co, cin, kh, kw] = sum_{b,i,j} dy[b,co,i,j] * x[b,cin,i*s+kh*dil-pad, j*s+kw*dil-pad]
struct ConvBwdWeightP {
  int M, N, K, k_chunk;
  const float* dy; const float* x; float* partial;
  int Cin, Hin, Win, Ho, Wo, ks, stride, pad, dil, Ctot;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
  struct ARow { const float* p; };
  struct BCol { const float* xc; int dh, dw; };
  __device__ ARow a_row(int m) const { return {dy + (size_t)m * Ho * Wo}; }
  __device__ float a(const ARow& r, int k) const {
    const int hw = Ho * Wo, b = k / hw, rem = k - b * hw;
    return __ldg(r.p + (size_t)b * Ctot * hw + rem);
  }
  __device__ BCol b_col(int n) const {
    const int kk = ks * ks, c = n / kk, rem = n - c * kk, kh = rem / ks, kw = rem - kh * ks;
    return {x + (size_t)c * Hin * Win, kh * dil - pad, kw * dil - pad};
  }
  __device__ float b(const BCol& c, int k) const {
    const int hw = Ho * Wo, b_ = k / hw, rem = k - b_ * hw, i = rem / Wo, j = rem - i * Wo;
    const int h = i * stride + c.dh, w_ = j * stride + c.dw;
    return (h >= 0 && h < Hin && w_ >= 0 && w_ < Win)
               ? __ldg(c.xc + (size_t)b_ * Cin * Hin * Win + (size_t)h * Win + w_) : 0.f;
  }
  __device__ void store(int m, int n, float v, int split) const {
    partial[(size_t)split * M * N + (size_t)m * N + n] = v;
  }
};

int f32_conv_bwd(const Dims& d, const float* x, const float* w, const float* dy, float* dx, float* dw,
                 float* partial, cudaStream_t st) {
  if (d.Cc == 0) return 0;
  if (dx) {
    for (int rh = 0; rh < d.stride; ++rh)
      for (int rw = 0; rw < d.stride; ++rw) {
        ConvBwdDataP p;
        p.nkh = p.nkw = 0;
        for (int kh = 0; kh < d.ks; ++kh)
          if (((rh + d.pad - kh * d.dil) % d.stride + d.stride) % d.stride == 0) p.khs[p.nkh++] = kh;
        for (int kw = 0; kw < d.ks; ++kw)
          if (((rw + d.pad - kw * d.dil) % d.stride + d.stride) % d.stride == 0) p.kws[p.nkw++] = kw;
        p.Hc = d.Hin > rh ? (d.Hin - rh + d.stride - 1) / d.stride : 0;
        p.Wc = d.Win > rw ? (d.Win - rw + d.stride - 1) / d.stride : 0;
        p.M = d.B * p.Hc * p.Wc; p.N = d.Cin; p.K = d.Cc * p.nkh * p.nkw; p.k_chunk = p.K > 0 ? p.K : 1;
        p.dy = dy; p.w = w; p.dx = dx;
        p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.Ho = d.H; p.Wo = d.W;
        p.ks = d.ks; p.stride = d.stride; p.pad = d.pad; p.dil = d.dil; p.Ctot = d.Cout; p.Cc = d.Cc;
        p.rh = rh; p.rw = rw;
        // K == 0 (no tap reaches this class) still launches: the kernel then writes zeros.
        AACONV_TRY(launch_simt_gemm(p, 1, st, "conv_bwd_data_f32"));
      }
  }
  if (dw) {
    ConvBwdWeightP p;
    p.M = d.Cc; p.N = d.Cin * d.ks * d.ks; p.K = d.B * d.L;
    int splits = pick_splits(p.M, p.N, p.K);
    p.k_chunk = chunk_for(p.K, splits);
    splits = cdiv(p.K, p.k_chunk);
    p.dy = dy; p.x = x; p.partial = partial;
    p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.Ho = d.H; p.Wo = d.W;
    p.ks = d.ks; p.stride = d.stride; p.pad = d.pad; p.dil = d.dil; p.Ctot = d.Cout;
    AACONV_TRY(launch_simt_gemm(p, splits, st, "conv_bwd_weight_f32"));
    AACONV_TRY(splitk_reduce(partial, dw, p.M * p.N, splits, st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// d key_rel[d, r] = sum_rows q[row, d] * dR[row, r]               (SURVEY.md 8a backward contract)
// ------------------------------------------------------------------------------------------------
struct RelWeightGradP {
  int M, N, K, k_chunk;
  const float* q; const float* dr; float* partial;
  int dkh, R;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = false;
  struct ARow { const float* p; };
  struct BCol { const float* p; };
  __device__ ARow a_row(int m) const { return {q + m}; }
  __device__ float a(const ARow& r, int k) const { return __ldg(r.p + (size_t)k * dkh); }
  __device__ BCol b_col(int n) const { return {dr + n}; }
  __device__ float b(const BCol& c, int k) const { return __ldg(c.p + (size_t)k * R); }
  __device__ void store(int m, int n, float v, int split) const {
    partial[(size_t)split * M * N + (size_t)m * N + n] = v;
  }
};

int f32_rel_weight_grad(const Dims& d, const float* q, const float* dr, int R, float* dkr, float* partial,
                        cudaStream_t st) {
  RelWeightGradP p;
  p.M = d.dkh; p.N = R; p.K = d.BN * d.L;
  int splits = pick_splits(p.M, p.N, p.K);
  p.k_chunk = chunk_for(p.K, splits);
  splits = cdiv(p.K, p.k_chunk);
  p.q = q; p.dr = dr; p.partial = partial; p.dkh = d.dkh; p.R = R;
  AACONV_TRY(launch_simt_gemm(p, splits, st, "rel_weight_grad_f32"));
  return splitk_reduce(partial, dkr, p.M * p.N, splits, st);
}

// d key_rel[e, r] from the abs-indexed gradients of the augmented rows (bf16 path):
//   axis 0 (W): dkr[e, r] = sum_rows q[row, e] * dQa[row, dkh + (r - (W-1) + x_row)]   (term absent when out of range)
//   axis 1 (H): same with y_row and the H block.
struct AugRelWeightGradP {
  int M, N, K, k_chunk;
  const float* q; const float* dqa; float* partial;
  int dkh, KD, L, W, n_axis, col0, is_w;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
  struct ARow { const float* p; };
  struct BCol { int shift; };
  __device__ ARow a_row(int m) const { return {q + m}; }
  __device__ float a(const ARow& r, int k) const { return __ldg(r.p + (size_t)k * dkh); }
  __device__ BCol b_col(int n) const { return {n - (n_axis - 1)}; }
  __device__ float b(const BCol& c, int k) const {
    const int l = k % L, y = l / W, x = l - y * W;
    const int pos = c.shift + (is_w ? x : y);
    return (pos >= 0 && pos < n_axis) ? __ldg(dqa + (size_t)k * KD + col0 + pos) : 0.f;
  }
  __device__ void store(int m, int n, float v, int split) const {
    partial[(size_t)split * M * N + (size_t)m * N + n] = v;
  }
};

int aug_rel_weight_grad(const Dims& d, const float* q, const float* dqa, int KD, int axis, float* dkr, float* partial,
                        cudaStream_t st) {
  AugRelWeightGradP p;
  p.M = d.dkh; p.N = axis ? d.RH : d.RW; p.K = d.BN * d.L;
  int splits = pick_splits(p.M, p.N, p.K);
  p.k_chunk = chunk_for(p.K, splits);
  splits = cdiv(p.K, p.k_chunk);
  p.q = q; p.dqa = dqa; p.partial = partial; p.dkh = d.dkh; p.KD = KD; p.L = d.L; p.W = d.W;
  p.n_axis = axis ? d.H : d.W; p.col0 = axis ? d.dkh + d.W : d.dkh; p.is_w = axis ? 0 : 1;
  AACONV_TRY(launch_simt_gemm(p, splits, st, "aug_rel_weight_grad"));
  return splitk_reduce(partial, dkr, p.M * p.N, splits, st);
}

// upper bound (floats) of the split-K partial buffer any of the launches above may need
size_t f32_partial_floats(const Dims& d) {
  auto need = [](int M, int N, int K) {
    int s = pick_splits(M, N, K);
    int c = chunk_for(K, s);
    s = cdiv(K, c);
    return (size_t)s * M * N;
  };
  size_t n = need(d.dv, d.dv, d.B * d.L);
  n = std::max(n, need(d.Nqkv, d.Cin, d.B * d.L));
  if (d.Cc) n = std::max(n, need(d.Cc, d.Cin * d.ks * d.ks, d.B * d.L));
  if (d.relative) {
    n = std::max(n, need(d.dkh, d.RW, d.BN * d.L));
    n = std::max(n, need(d.dkh, d.RH, d.BN * d.L));
  }
  return n;
}

}  // namespace aaconv
