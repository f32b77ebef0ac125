// fp32 (FFMA) attention core of AAConv2d: flash-style, nothing of size (HW x HW) is materialised
// unless the caller asks for the attention map.  Reference rows a3-a8 (attn_aug_conv.py:75-91):
//   logit[q,(y',x')] = q.k + Rw[q, x'-x+W-1] + Rh[q, y'-y+H-1],  Rw = q.key_rel_w, Rh = q.key_rel_h
// The rel_to_abs pad/reshape of attn_aug_conv.py:43-53 is replaced by that index computation.
#include "fp32_path.cuh"

namespace aaconv {

constexpr int TQ = 128;   // threads per block == rows (queries or keys) per block
constexpr int KT = 64;    // keys per shared-memory tile (forward / dq pass)
constexpr int QT = 32;    // queries per shared-memory tile (dk/dv pass)

template <int DK>
__device__ __forceinline__ float dot_smem(const float* __restrict__ row, const float (&r)[DK]) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < DK; d += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + d);
    s = fmaf(r[d], t.x, s); s = fmaf(r[d + 1], t.y, s); s = fmaf(r[d + 2], t.z, s); s = fmaf(r[d + 3], t.w, s);
  }
  return s;
}

// rows of a (rows, dh) global tensor -> zero-padded (tile, DP) shared tile
template <int DP>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ src, int n_valid,
                                          int n_tile, int dh) {
  for (int idx = threadIdx.x; idx < n_tile * DP; idx += blockDim.x) {
    const int r = idx / DP, e = idx - r * DP;
    dst[idx] = (r < n_valid && e < dh) ? __ldg(src + (size_t)r * dh + e) : 0.f;
  }
}

// ---- Rw / Rh tables -----------------------------------------------------------------------------
__global__ void f32_rel_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kr,
                                   float* __restrict__ out, size_t rows, int dkh, int R) {
  extern __shared__ float krs[];
  for (int i = threadIdx.x; i < dkh * R; i += blockDim.x) krs[i] = kr[i];
  __syncthreads();
  const size_t total = rows * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / R;
    const int r = (int)(i - row * R);
    float s = 0.f;
    for (int d = 0; d < dkh; ++d) s = fmaf(__ldg(q + row * dkh + d), krs[d * R + r], s);
    out[i] = s;
  }
}

int f32_rel_fwd(const Dims& d, const float* q, const float* krw, const float* krh, float* rw, float* rh,
                cudaStream_t st) {
  if (!d.relative) return 0;
  const size_t rows = (size_t)d.BN * d.L;
  for (int axis = 0; axis < 2; ++axis) {
    const int R = axis ? d.RH : d.RW;
    const size_t smem = (size_t)d.dkh * R * sizeof(float);
    if (smem > 200 * 1024) return fail(AACONV_E_UNSUPPORTED, "key_rel table too large for shared memory");
    AACONV_CUDA_OK(cudaFuncSetAttribute(f32_rel_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<size_t>((rows * R + 255) / 256, 148 * 16);
    f32_rel_fwd_kernel<<<grid, 256, smem, AACONV_ST(st)>>>(q, axis ? krh : krw, axis ? rh : rw, rows, d.dkh, R);
    AACONV_LAUNCH_OK("rel_fwd_f32");
  }
  return 0;
}

// ---- forward (online softmax) and attention-map writer ------------------------------------------
template <int DK, int DV, bool WRITE_P>
__global__ void __launch_bounds__(TQ) f32_attn_fwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ rw, const float* __restrict__ rh, float* __restrict__ o,
    float* __restrict__ lse, float* __restrict__ weights, int L, int H, int W, int dkh, int dvh, int relative) {
  extern __shared__ __align__(16) float smem[];
  float* ks = smem;
  float* vs = ks + KT * DK;
  float* rws = vs + KT * DV;
  const int tid = threadIdx.x, bn = blockIdx.y;
  const int qraw = blockIdx.x * TQ + tid;
  const bool valid = qraw < L;
  const int qi = valid ? qraw : L - 1;
  const int qy = qi / W, qx = qi - qy * W;
  const int RW = 2 * W - 1, RH = 2 * H - 1;
  const size_t row = (size_t)bn * L + qi;

  float qr[DK];
#pragma unroll
  for (int e = 0; e < DK; ++e) qr[e] = e < dkh ? __ldg(q + row * dkh + e) : 0.f;
  if (relative)
    for (int r = 0; r < RW; ++r) rws[r * TQ + tid] = __ldg(rw + row * RW + r);
  const float* rh_row = relative ? rh + row * RH + (H - 1 - qy) : nullptr;   // indexed by key row y'
  const float* rw_col = rws + (W - 1 - qx) * TQ + tid;                       // indexed by key col x' (stride TQ)

  float m = -INFINITY, l = 0.f, acc[DV];
#pragma unroll
  for (int e = 0; e < DV; ++e) acc[e] = 0.f;
  const float lse_q = WRITE_P ? __ldg(lse + row) : 0.f;
  int ky = 0, kx = 0;
  float rhv = relative ? __ldg(rh_row) : 0.f;

  for (int k0 = 0; k0 < L; k0 += KT) {
    __syncthreads();
    const int kn = min(KT, L - k0);
    load_tile<DK>(ks, k + ((size_t)bn * L + k0) * dkh, kn, KT, dkh);
    if (!WRITE_P) load_tile<DV>(vs, v + ((size_t)bn * L + k0) * dvh, kn, KT, dvh);
    __syncthreads();
    for (int kk = 0; kk < kn; ++kk) {
      float s = dot_smem<DK>(ks + kk * DK, qr);
      if (relative) s += rw_col[kx * TQ] + rhv;
      if (WRITE_P) {
        if (valid) weights[row * L + k0 + kk] = expf(s - lse_q);
      } else {
        if (s > m) {                      // rescale the running sums to the new maximum
          const float c = expf(m - s);    // m == -inf -> 0
          l *= c;
#pragma unroll
          for (int e = 0; e < DV; ++e) acc[e] *= c;
          m = s;
        }
        const float p = expf(s - m);
        l += p;
#pragma unroll
        for (int e = 0; e < DV; ++e) acc[e] = fmaf(p, vs[kk * DV + e], acc[e]);
      }
      if (++kx == W) {
        kx = 0;
        ++ky;
        if (relative && ky < H) rhv = __ldg(rh_row + ky);
      }
    }
  }
  if (!WRITE_P && valid) {
    const float inv = 1.f / l;
#pragma unroll
    for (int e = 0; e < DV; ++e)
      if (e < dvh) o[row * dvh + e] = acc[e] * inv;
    lse[row] = m + logf(l);
  }
}

// ---- backward, query-stationary pass: dq (content part), dRw, dRh ---------------------------------
template <int DK, int DV>
__global__ void __launch_bounds__(TQ) f32_attn_bwd_dq_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ rw, const float* __restrict__ rh, const float* __restrict__ lse,
    const float* __restrict__ d_o, const float* __restrict__ delta, float* __restrict__ dq,
    float* __restrict__ drw, float* __restrict__ drh, int L, int H, int W, int dkh, int dvh, int relative) {
  extern __shared__ __align__(16) float smem[];
  float* ks = smem;
  float* vs = ks + KT * DK;
  float* rws = vs + KT * DV;
  const int RW = 2 * W - 1, RH = 2 * H - 1;
  float* drws = rws + (relative ? RW * TQ : 0);
  const int tid = threadIdx.x, bn = blockIdx.y;
  const int qraw = blockIdx.x * TQ + tid;
  const bool valid = qraw < L;
  const int qi = valid ? qraw : L - 1;
  const int qy = qi / W, qx = qi - qy * W;
  const size_t row = (size_t)bn * L + qi;

  float qr[DK], dqr[DK], dor[DV];
#pragma unroll
  for (int e = 0; e < DK; ++e) { qr[e] = e < dkh ? __ldg(q + row * dkh + e) : 0.f; dqr[e] = 0.f; }
#pragma unroll
  for (int e = 0; e < DV; ++e) dor[e] = e < dvh ? __ldg(d_o + row * dvh + e) : 0.f;
  if (relative)
    for (int r = 0; r < RW; ++r) { rws[r * TQ + tid] = __ldg(rw + row * RW + r); drws[r * TQ + tid] = 0.f; }
  const float* rh_row = relative ? rh + row * RH + (H - 1 - qy) : nullptr;
  float* drh_row = relative ? drh + row * RH + (H - 1 - qy) : nullptr;
  const int col0 = (W - 1 - qx) * TQ + tid;
  const float lse_q = __ldg(lse + row), delta_q = __ldg(delta + row);
  int ky = 0, kx = 0;
  float rhv = relative ? __ldg(rh_row) : 0.f, drh_acc = 0.f;

  for (int k0 = 0; k0 < L; k0 += KT) {
    __syncthreads();
    const int kn = min(KT, L - k0);
    load_tile<DK>(ks, k + ((size_t)bn * L + k0) * dkh, kn, KT, dkh);
    load_tile<DV>(vs, v + ((size_t)bn * L + k0) * dvh, kn, KT, dvh);
    __syncthreads();
    for (int kk = 0; kk < kn; ++kk) {
      float s = dot_smem<DK>(ks + kk * DK, qr);
      if (relative) s += rws[col0 + kx * TQ] + rhv;
      const float p = expf(s - lse_q);
      float dp = 0.f;
#pragma unroll
      for (int e = 0; e < DV; ++e) dp = fmaf(dor[e], vs[kk * DV + e], dp);
      const float ds = p * (dp - delta_q);
#pragma unroll
      for (int e = 0; e < DK; ++e) dqr[e] = fmaf(ds, ks[kk * DK + e], dqr[e]);
      if (relative) { drws[col0 + kx * TQ] += ds; drh_acc += ds; }
      if (++kx == W) {
        if (relative && valid) drh_row[ky] = drh_acc;   // each (query, key row) pair maps to one slot
        drh_acc = 0.f;
        kx = 0;
        ++ky;
        if (relative && ky < H) rhv = __ldg(rh_row + ky);
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int e = 0; e < DK; ++e)
      if (e < dkh) dq[row * dkh + e] = dqr[e];
    if (relative) {
      // slots the query can never reach (x'-x+W-1 outside [W-1-x, 2W-2-x]) stay exactly zero
      for (int r = 0; r < RW; ++r) drw[row * RW + r] = drws[r * TQ + tid];
      for (int r = 0; r < H - 1 - qy; ++r) drh[row * RH + r] = 0.f;
      for (int r = 2 * H - 1 - qy; r < RH; ++r) drh[row * RH + r] = 0.f;
    }
  }
}

// ---- backward, key-stationary pass: dk, dv ---------------------------------------------------------
template <int DK, int DV>
__global__ void __launch_bounds__(TQ) f32_attn_bwd_dkv_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ rw, const float* __restrict__ rh, const float* __restrict__ lse,
    const float* __restrict__ d_o, const float* __restrict__ delta, float* __restrict__ dk,
    float* __restrict__ dv, int L, int H, int W, int dkh, int dvh, int relative) {
  extern __shared__ __align__(16) float smem[];
  const int RW = 2 * W - 1, RH = 2 * H - 1;
  float* qs = smem;
  float* dos = qs + QT * DK;
  float* lses = dos + QT * DV;
  float* deltas = lses + QT;
  float* rwt = deltas + QT;
  float* rht = rwt + (relative ? QT * RW : 0);
  const int tid = threadIdx.x, bn = blockIdx.y;
  const int kraw = blockIdx.x * TQ + tid;
  const bool valid = kraw < L;
  const int kj = valid ? kraw : L - 1;
  const int ky = kj / W, kx = kj - ky * W;
  const size_t krow = (size_t)bn * L + kj;

  float kr[DK], dkr[DK], vr[DV], dvr[DV];
#pragma unroll
  for (int e = 0; e < DK; ++e) { kr[e] = e < dkh ? __ldg(k + krow * dkh + e) : 0.f; dkr[e] = 0.f; }
#pragma unroll
  for (int e = 0; e < DV; ++e) { vr[e] = e < dvh ? __ldg(v + krow * dvh + e) : 0.f; dvr[e] = 0.f; }

  int qy = 0, qx = 0;
  for (int q0 = 0; q0 < L; q0 += QT) {
    __syncthreads();
    const int qn = min(QT, L - q0);
    const size_t r0 = (size_t)bn * L + q0;
    load_tile<DK>(qs, q + r0 * dkh, qn, QT, dkh);
    load_tile<DV>(dos, d_o + r0 * dvh, qn, QT, dvh);
    for (int i = tid; i < qn; i += TQ) { lses[i] = __ldg(lse + r0 + i); deltas[i] = __ldg(delta + r0 + i); }
    if (relative) {
      for (int i = tid; i < qn * RW; i += TQ) rwt[i] = __ldg(rw + r0 * RW + i);
      for (int i = tid; i < qn * RH; i += TQ) rht[i] = __ldg(rh + r0 * RH + i);
    }
    __syncthreads();
    for (int qq = 0; qq < qn; ++qq) {
      float s = dot_smem<DK>(qs + qq * DK, kr);
      if (relative) s += rwt[qq * RW + kx - qx + W - 1] + rht[qq * RH + ky - qy + H - 1];
      const float p = expf(s - lses[qq]);
      float dp = 0.f;
#pragma unroll
      for (int e = 0; e < DV; ++e) {
        const float g = dos[qq * DV + e];
        dvr[e] = fmaf(p, g, dvr[e]);
        dp = fmaf(g, vr[e], dp);
      }
      const float ds = p * (dp - deltas[qq]);
#pragma unroll
      for (int e = 0; e < DK; ++e) dkr[e] = fmaf(ds, qs[qq * DK + e], dkr[e]);
      if (++qx == W) { qx = 0; ++qy; }
    }
  }
  if (valid) {
#pragma unroll
    for (int e = 0; e < DK; ++e)
      if (e < dkh) dk[krow * dkh + e] = dkr[e];
#pragma unroll
    for (int e = 0; e < DV; ++e)
      if (e < dvh) dv[krow * dvh + e] = dvr[e];
  }
}

// ---- small element-wise helpers --------------------------------------------------------------------
__global__ void f32_delta_kernel(const float* __restrict__ d_o, const float* __restrict__ o,
                                 float* __restrict__ delta, size_t rows, int dvh) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float s = 0.f;
  for (int e = 0; e < dvh; ++e) s = fmaf(d_o[i * dvh + e], o[i * dvh + e], s);
  delta[i] = s;
}

int f32_delta(const Dims& d, const float* d_o, const float* o, float* delta, cudaStream_t st) {
  const size_t rows = (size_t)d.BN * d.L;
  f32_delta_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, AACONV_ST(st)>>>(d_o, o, delta, rows, d.dvh);
  AACONV_LAUNCH_OK("delta_f32");
  return 0;
}

// dq[row, e] += sum_r key_rel_w[e, r] dRw[row, r] + sum_r key_rel_h[e, r] dRh[row, r]
__global__ void f32_rel_bwd_dq_kernel(const float* __restrict__ krw, const float* __restrict__ krh,
                                      const float* __restrict__ drw, const float* __restrict__ drh,
                                      float* __restrict__ dq, size_t rows, int dkh, int RW, int RH) {
  extern __shared__ float sm[];
  float* kw = sm;
  float* kh = sm + dkh * RW;
  for (int i = threadIdx.x; i < dkh * RW; i += blockDim.x) kw[i] = krw[i];
  for (int i = threadIdx.x; i < dkh * RH; i += blockDim.x) kh[i] = krh[i];
  __syncthreads();
  const size_t total = rows * dkh;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / dkh;
    const int e = (int)(i - row * dkh);
    float s = 0.f;
    for (int r = 0; r < RW; ++r) s = fmaf(kw[e * RW + r], __ldg(drw + row * RW + r), s);
    for (int r = 0; r < RH; ++r) s = fmaf(kh[e * RH + r], __ldg(drh + row * RH + r), s);
    dq[i] += s;
  }
}

int f32_rel_bwd_dq(const Dims& d, const float* krw, const float* krh, const float* drw, const float* drh,
                   float* dq, cudaStream_t st) {
  if (!d.relative) return 0;
  const size_t rows = (size_t)d.BN * d.L;
  const size_t smem = (size_t)d.dkh * (d.RW + d.RH) * sizeof(float);
  if (smem > 200 * 1024) return fail(AACONV_E_UNSUPPORTED, "key_rel tables too large for shared memory");
  AACONV_CUDA_OK(cudaFuncSetAttribute(f32_rel_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<size_t>((rows * d.dkh + 255) / 256, 148 * 16);
  f32_rel_bwd_dq_kernel<<<grid, 256, smem, AACONV_ST(st)>>>(krw, krh, drw, drh, dq, rows, d.dkh, d.RW, d.RH);
  AACONV_LAUNCH_OK("rel_bwd_dq_f32");
  return 0;
}

// ---- launchers with (dkh, dvh) -> padded template dispatch -----------------------------------------
#define AACONV_DISPATCH_DV(DKP, ...)                       \
  if (d.dvh <= 1) { constexpr int DK = DKP, DV = 1; __VA_ARGS__ }  \
  else if (d.dvh <= 4) { constexpr int DK = DKP, DV = 4; __VA_ARGS__ }  \
  else if (d.dvh <= 8) { constexpr int DK = DKP, DV = 8; __VA_ARGS__ }  \
  else { constexpr int DK = DKP, DV = 64; __VA_ARGS__ }

#define AACONV_DISPATCH(...)                                                                       \
  do {                                                                                             \
    if (d.dkh > 64 || d.dvh > 64)                                                                  \
      return fail(AACONV_E_UNSUPPORTED, "fp32 path supports dk/nh <= 64 and dv/nh <= 64 (got %d, %d)", d.dkh, d.dvh); \
    if (d.dkh <= 8) { AACONV_DISPATCH_DV(8, __VA_ARGS__) }                                                \
    else if (d.dkh <= 20) { AACONV_DISPATCH_DV(20, __VA_ARGS__) }                                         \
    else if (d.dkh <= 32) { AACONV_DISPATCH_DV(32, __VA_ARGS__) }                                         \
    else { AACONV_DISPATCH_DV(64, __VA_ARGS__) }                                                          \
  } while (0)

template <class K>
static int set_smem(K kernel, size_t smem) {
  if (smem > 227 * 1024) return fail(AACONV_E_UNSUPPORTED, "attention tile needs %zu B of shared memory (> 227 KB)", smem);
  AACONV_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

int f32_attn_fwd(const Dims& d, const float* q, const float* k, const float* v, const float* rw,
                 const float* rh, float* o, float* lse, cudaStream_t st) {
  dim3 grid(cdiv(d.L, TQ), d.BN);
  AACONV_DISPATCH({
    const size_t smem = sizeof(float) * ((size_t)KT * (DK + DV) + (d.relative ? (size_t)d.RW * TQ : 0));
    auto kern = f32_attn_fwd_kernel<DK, DV, false>;
    AACONV_TRY(set_smem(kern, smem));
    kern<<<grid, TQ, smem, AACONV_ST(st)>>>(q, k, v, rw, rh, o, lse, nullptr, d.L, d.H, d.W, d.dkh, d.dvh, d.relative);
  });
  AACONV_LAUNCH_OK("attn_fwd_f32");
  return 0;
}

int f32_attn_weights(const Dims& d, const float* q, const float* k, const float* rw, const float* rh,
                     const float* lse, float* weights, cudaStream_t st) {
  dim3 grid(cdiv(d.L, TQ), d.BN);
  AACONV_DISPATCH({
    const size_t smem = sizeof(float) * ((size_t)KT * (DK + DV) + (d.relative ? (size_t)d.RW * TQ : 0));
    auto kern = f32_attn_fwd_kernel<DK, DV, true>;
    AACONV_TRY(set_smem(kern, smem));
    kern<<<grid, TQ, smem, AACONV_ST(st)>>>(q, k, nullptr, rw, rh, nullptr, const_cast<float*>(lse), weights, d.L, d.H, d.W,
                                 d.dkh, d.dvh, d.relative);
  });
  AACONV_LAUNCH_OK("attn_weights_f32");
  return 0;
}

int f32_attn_bwd(const Dims& d, const float* q, const float* k, const float* v, const float* rw,
                 const float* rh, const float* lse, const float* d_o, const float* delta, float* dq,
                 float* dk, float* dv, float* drw, float* drh, cudaStream_t st) {
  dim3 grid(cdiv(d.L, TQ), d.BN);
  AACONV_DISPATCH({
    const size_t smem = sizeof(float) * ((size_t)KT * (DK + DV) + (d.relative ? 2 * (size_t)d.RW * TQ : 0));
    auto kern = f32_attn_bwd_dq_kernel<DK, DV>;
    AACONV_TRY(set_smem(kern, smem));
    kern<<<grid, TQ, smem, AACONV_ST(st)>>>(q, k, v, rw, rh, lse, d_o, delta, dq, drw, drh, d.L, d.H, d.W, d.dkh, d.dvh,
                                 d.relative);
  });
  AACONV_LAUNCH_OK("attn_bwd_dq_f32");
  AACONV_DISPATCH({
    const size_t smem = sizeof(float) * ((size_t)QT * (DK + DV + 2) + (d.relative ? (size_t)QT * (d.RW + d.RH) : 0));
    auto kern = f32_attn_bwd_dkv_kernel<DK, DV>;
    AACONV_TRY(set_smem(kern, smem));
    kern<<<grid, TQ, smem, AACONV_ST(st)>>>(q, k, v, rw, rh, lse, d_o, delta, dk, dv, d.L, d.H, d.W, d.dkh, d.dvh, d.relative);
  });
  AACONV_LAUNCH_OK("attn_bwd_dkv_f32");
  return 0;
}

}  // namespace aaconv
