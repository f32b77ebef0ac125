// bf16 tcgen05 attention core of AAConv2d (forward): TMA-fed QK^T into TMEM, online softmax from TMEM,
// P.V from TMEM-resident P.  Reference rows a3-a8 (attn_aug_conv.py:75-91).
//
// Relative logits as part of the QK^T contraction ("augmented K dimension"):
//   logit[q,(y',x')] = q.k + Rw[q, x'-x+W-1] + Rh[q, y'-y+H-1]
//                    = [ q | Aq | Bq ] . [ k | onehot_W(x') | onehot_H(y') ]
//   Aq[x'] = q.key_rel_w[:, x'-x+W-1]   (the index computation that replaces rel_to_abs, attn_aug_conv.py:43-53)
//   Bq[y'] = q.key_rel_h[:, y'-y+H-1]
// so one MMA over K = dkh + W + H produces content + both relative terms; nothing of size (HW x HW)
// ever leaves TMEM.  A "ones" row appended to V makes the same P.V MMA produce the softmax row sum.
#include "tc_common.cuh"
#include "bf16_path.cuh"

namespace aaconv {

using tc::smem_u32;
typedef __nv_bfloat16 bf16;

constexpr int FA_BM = 128;      // queries per CTA  (TMEM lanes)
constexpr int FA_BN = 128;      // keys per tile
constexpr int FA_STAGES = 2;    // K/V smem stages
constexpr int FA_DV = 16;       // padded value width: dvh values, then the ones column, then zeros
constexpr int FA_THREADS = 192; // warps 0-3 softmax, warp 4 TMA, warp 5 MMA (+TMEM alloc)
constexpr uint32_t FA_TMEM_COLS = 256;
constexpr uint32_t FA_COL_S = 0, FA_COL_P = 128, FA_COL_O = 192;

template <int KATOMS>
struct __align__(1024) FwdSmem {
  bf16 q[KATOMS][FA_BM * 64];                 // K-major, 128 B rows, 128B swizzle (one atom = 64 k-elements)
  bf16 k[FA_STAGES][KATOMS][FA_BN * 64];
  bf16 v[FA_STAGES][2][FA_DV * 64];           // V^T tile: FA_DV rows x 64 keys per atom
  uint64_t bar_q, bar_kv_full[FA_STAGES], bar_kv_empty[FA_STAGES], bar_s_full, bar_p_ready, bar_o_done;
  uint32_t tmem_base;
};

template <int KATOMS>
__global__ void __launch_bounds__(FA_THREADS) attn_fwd_tc_kernel(
    const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
    const __grid_constant__ CUtensorMap tm_v, float* __restrict__ o, float* __restrict__ lse, int L, int dvh,
    int ksteps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FwdSmem<KATOMS>& sm = *reinterpret_cast<FwdSmem<KATOMS>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = blockIdx.y, q0 = blockIdx.x * FA_BM;
  const int ntiles = (L + FA_BN - 1) / FA_BN;

  if (threadIdx.x == 0) {
    tc::mbar_init(&sm.bar_q, 1);
    for (int s = 0; s < FA_STAGES; ++s) { tc::mbar_init(&sm.bar_kv_full[s], 1); tc::mbar_init(&sm.bar_kv_empty[s], 1); }
    tc::mbar_init(&sm.bar_s_full, 1);
    tc::mbar_init(&sm.bar_p_ready, 128);
    tc::mbar_init(&sm.bar_o_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 4 && lane == 0) { tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v); }
  if (warp == 5) tc::tmem_alloc<FA_TMEM_COLS>(&sm.tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&sm.bar_q, KATOMS * FA_BM * 64 * 2);
      for (int a = 0; a < KATOMS; ++a) tc::tma_load_3d(sm.q[a], &tm_q, &sm.bar_q, a * 64, q0, bn);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % FA_STAGES, ph = (j / FA_STAGES) & 1;
        tc::mbar_wait(&sm.bar_kv_empty[s], ph ^ 1);
        tc::mbar_arrive_expect_tx(&sm.bar_kv_full[s], KATOMS * FA_BN * 64 * 2 + 2 * FA_DV * 64 * 2);
        for (int a = 0; a < KATOMS; ++a) tc::tma_load_3d(sm.k[s][a], &tm_k, &sm.bar_kv_full[s], a * 64, j * FA_BN, bn);
        for (int a = 0; a < 2; ++a) tc::tma_load_2d(sm.v[s][a], &tm_v, &sm.bar_kv_full[s], j * FA_BN + a * 64, bn * FA_DV);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = tc::idesc_bf16_f32(FA_BM, FA_BN);
    constexpr uint32_t idesc_o = tc::idesc_bf16_f32(FA_BM, FA_DV);
    constexpr uint32_t Q_ATOM = (FA_BM * 128) >> 4, K_ATOM = (FA_BN * 128) >> 4, K_STAGE = KATOMS * K_ATOM;
    constexpr uint32_t V_ATOM = (FA_DV * 128) >> 4, V_STAGE = 2 * V_ATOM;
    const uint32_t q_lo = tc::desc_lo_k(smem_u32(sm.q[0]));
    const uint32_t k_lo = tc::desc_lo_k(smem_u32(sm.k[0][0]));
    const uint32_t v_lo = tc::desc_lo_k(smem_u32(sm.v[0][0]));
    tc::mbar_wait(&sm.bar_q, 0);
    for (int j = 0; j < ntiles; ++j) {
      const int s = j % FA_STAGES, ph = (j / FA_STAGES) & 1;
      tc::mbar_wait(&sm.bar_kv_full[s], ph);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks)
          tc::mma_ss(tmem + FA_COL_S, tc::desc64(q_lo + (ks >> 2) * Q_ATOM + (ks & 3) * 2),
                     tc::desc64(k_lo + s * K_STAGE + (ks >> 2) * K_ATOM + (ks & 3) * 2), idesc_s, ks > 0);
        tc::mma_commit(&sm.bar_s_full);
      }
      __syncwarp();
      tc::mbar_wait(&sm.bar_p_ready, j & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < FA_BN / 16; ++ks)
          tc::mma_ts(tmem + FA_COL_O, tmem + FA_COL_P + ks * 8,
                     tc::desc64(v_lo + s * V_STAGE + (ks >> 2) * V_ATOM + (ks & 3) * 2), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        tc::mma_commit(&sm.bar_kv_empty[s]);
        tc::mma_commit(&sm.bar_o_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax warps (thread == query row == TMEM lane) =====================
    const float LOG2E = 1.4426950408889634f;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    float m = -INFINITY;
    uint32_t r[32];
    for (int j = 0; j < ntiles; ++j) {
      const int nvalid = min(FA_BN, L - j * FA_BN);
      tc::mbar_wait(&sm.bar_s_full, j & 1);
      tc::tc_fence_after();
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < FA_BN / 32; ++c) {
        tc::tmem_ld_x32(tlane + FA_COL_S + c * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = __uint_as_float(r[i]);
          tmax = fmaxf(tmax, (c * 32 + i < nvalid) ? sv : -INFINITY);
        }
      }
      const float m_new = fmaxf(m, tmax);
      if (j > 0) {
        tc::mbar_wait(&sm.bar_o_done, (j - 1) & 1);      // P.V of the previous tile has consumed P and updated O
        tc::tc_fence_after();
        if (__any_sync(0xffffffffu, m_new > m)) {
          const float alpha = tc::ex2f((m - m_new) * LOG2E);
          uint32_t ov[16];
          tc::tmem_ld_x16(tlane + FA_COL_O, ov);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tc::tmem_st_x16(tlane + FA_COL_O, ov);
        }
      }
      m = m_new;
      const float mneg = -m * LOG2E;
#pragma unroll
      for (int c = 0; c < FA_BN / 32; ++c) {
        tc::tmem_ld_x32(tlane + FA_COL_S + c * 32, r);
        tc::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = tc::ex2f(fmaf(__uint_as_float(r[2 * i]), LOG2E, mneg));
          float p1 = tc::ex2f(fmaf(__uint_as_float(r[2 * i + 1]), LOG2E, mneg));
          if (c * 32 + 2 * i >= nvalid) p0 = 0.f;
          if (c * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
          pk[i] = tc::pack_bf16x2(p0, p1);
        }
        tc::tmem_st_x16(tlane + FA_COL_P + c * 16, pk);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&sm.bar_p_ready);
    }
    tc::mbar_wait(&sm.bar_o_done, (ntiles - 1) & 1);
    tc::tc_fence_after();
    uint32_t ov[16];
    tc::tmem_ld_x16(tlane + FA_COL_O, ov);
    tc::tmem_ld_wait();
    const int qi = q0 + threadIdx.x;
    if (qi < L) {
      float l = 1.f;                                     // the ones column of V: softmax row sum
#pragma unroll
      for (int e = 0; e < FA_DV; ++e)
        if (e == dvh) l = __uint_as_float(ov[e]);
      const float inv = 1.f / l;
      const size_t row = (size_t)bn * L + qi;
#pragma unroll
      for (int e = 0; e < FA_DV; ++e)
        if (e < dvh) o[row * dvh + e] = __uint_as_float(ov[e]) * inv;
      lse[row] = m + logf(l);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc<FA_TMEM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------------
// operand builders: fp32 head-split q,k,v -> augmented bf16 operands
// ------------------------------------------------------------------------------------------------
// qa[row, c], ka[row, c] for c < KP (KP = 64 * KATOMS); layout (BN, L, KP)
__global__ void attn_aug_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                const float* __restrict__ krw, const float* __restrict__ krh,
                                bf16* __restrict__ qa, bf16* __restrict__ ka, size_t rows, int L, int H, int W, int dkh,
                                int KP, int relative) {
  extern __shared__ float sm[];
  const int RW = 2 * W - 1, RH = 2 * H - 1;
  float* kw = sm;                                   // dkh x RW
  float* kh = kw + (relative ? dkh * RW : 0);       // dkh x RH
  float* qs = kh + (relative ? dkh * RH : 0);       // 16 x dkh
  float* ks = qs + 16 * dkh;
  if (relative) {
    for (int i = threadIdx.x; i < dkh * RW; i += blockDim.x) kw[i] = krw[i];
    for (int i = threadIdx.x; i < dkh * RH; i += blockDim.x) kh[i] = krh[i];
  }
  const int c = threadIdx.x;                        // blockDim.x == KP
  for (size_t r0 = (size_t)blockIdx.x * 16; r0 < rows; r0 += (size_t)gridDim.x * 16) {
    __syncthreads();
    const int nr = (int)min((size_t)16, rows - r0);
    for (int i = threadIdx.x; i < nr * dkh; i += blockDim.x) { qs[i] = q[r0 * dkh + i]; ks[i] = k[r0 * dkh + i]; }
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
      const size_t row = r0 + rr;
      const int l = (int)(row % L), y = l / W, x = l - y * W;
      float qv = 0.f, kv = 0.f;
      if (c < dkh) {
        qv = qs[rr * dkh + c];
        kv = ks[rr * dkh + c];
      } else if (relative && c < dkh + W) {
        const int xp = c - dkh, rix = xp - x + W - 1;
        for (int e = 0; e < dkh; ++e) qv = fmaf(qs[rr * dkh + e], kw[e * RW + rix], qv);
        kv = (xp == x) ? 1.f : 0.f;
      } else if (relative && c < dkh + W + H) {
        const int yp = c - dkh - W, riy = yp - y + H - 1;
        for (int e = 0; e < dkh; ++e) qv = fmaf(qs[rr * dkh + e], kh[e * RH + riy], qv);
        kv = (yp == y) ? 1.f : 0.f;
      }
      qa[row * KP + c] = __float2bfloat16(qv);
      ka[row * KP + c] = __float2bfloat16(kv);
    }
  }
}

// vt[bn, e, l] (FA_DV rows, row stride Lp): e < dvh -> v, e == dvh -> 1, else 0
__global__ void attn_vt_kernel(const float* __restrict__ v, bf16* __restrict__ vt, int BN, int L, int Lp, int dvh) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)BN * Lp) return;
  const int bn = (int)(i / Lp), l = (int)(i - (size_t)bn * Lp);
  for (int e = 0; e < FA_DV; ++e) {
    float val = 0.f;
    if (l < L) val = e < dvh ? v[((size_t)bn * L + l) * dvh + e] : (e == dvh ? 1.f : 0.f);
    vt[((size_t)bn * FA_DV + e) * Lp + l] = __float2bfloat16(val);
  }
}

int tc_attn_kp(const Dims& d) {
  const int kdim = d.dkh + (d.relative ? d.W + d.H : 0);
  return cdiv(kdim, 64) * 64;
}

int tc_attn_supported(const Dims& d) {
  if (d.dvh + 1 > FA_DV) return fail(AACONV_E_UNSUPPORTED, "bf16 attention kernel supports dv/nh <= %d (got %d)", FA_DV - 1, d.dvh);
  if (tc_attn_kp(d) > 192)
    return fail(AACONV_E_UNSUPPORTED, "bf16 attention kernel supports dk/nh + H + W <= 192 (got %d)", d.dkh + d.W + d.H);
  return 0;
}

size_t tc_attn_operand_bytes(const Dims& d, size_t* qa_off, size_t* ka_off, size_t* vt_off) {
  const size_t rows = (size_t)d.BN * d.L;
  const int KP = tc_attn_kp(d), Lp = cdiv(d.L, 8) * 8;
  size_t off = 0;
  if (qa_off) *qa_off = off;
  off += align256(rows * KP * sizeof(bf16));
  if (ka_off) *ka_off = off;
  off += align256(rows * KP * sizeof(bf16));
  if (vt_off) *vt_off = off;
  off += align256((size_t)d.BN * FA_DV * Lp * sizeof(bf16));
  return off;
}

template <int KATOMS>
static int launch_fwd(const Dims& d, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, float* o,
                      float* lse, int ksteps, cudaStream_t st) {
  const size_t smem = sizeof(FwdSmem<KATOMS>) + 1024;
  auto kern = attn_fwd_tc_kernel<KATOMS>;
  AACONV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(d.L, FA_BM), d.BN);
  kern<<<grid, FA_THREADS, smem, st>>>(tq, tk, tv, o, lse, d.L, d.dvh, ksteps);
  AACONV_LAUNCH_OK("attn_fwd_tc");
  return 0;
}

// q,k,v: fp32 head-split (B,nh,L,dkh|dvh), q pre-scaled.  operands: scratch of tc_attn_operand_bytes().
int tc_attn_fwd(const Dims& d, const float* q, const float* k, const float* v, const float* krw, const float* krh,
                void* operands, float* o, float* lse, cudaStream_t st) {
  AACONV_TRY(tc_attn_supported(d));
  size_t qo, ko, vo;
  tc_attn_operand_bytes(d, &qo, &ko, &vo);
  bf16* qa = reinterpret_cast<bf16*>(static_cast<char*>(operands) + qo);
  bf16* ka = reinterpret_cast<bf16*>(static_cast<char*>(operands) + ko);
  bf16* vt = reinterpret_cast<bf16*>(static_cast<char*>(operands) + vo);
  const size_t rows = (size_t)d.BN * d.L;
  const int KP = tc_attn_kp(d), Lp = cdiv(d.L, 8) * 8;
  const int kdim = d.dkh + (d.relative ? d.W + d.H : 0);
  {
    const size_t smem = sizeof(float) * ((d.relative ? (size_t)d.dkh * (d.RW + d.RH) : 0) + 32 * (size_t)d.dkh);
    AACONV_CUDA_OK(cudaFuncSetAttribute(attn_aug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<size_t>((rows + 15) / 16, 148 * 8);
    attn_aug_kernel<<<grid, KP, smem, st>>>(q, k, krw, krh, qa, ka, rows, d.L, d.H, d.W, d.dkh, KP, d.relative);
    AACONV_LAUNCH_OK("attn_aug");
    attn_vt_kernel<<<(unsigned)(((size_t)d.BN * Lp + 255) / 256), 256, 0, st>>>(v, vt, d.BN, d.L, Lp, d.dvh);
    AACONV_LAUNCH_OK("attn_vt");
  }
  CUtensorMap tq, tk, tv;
  {
    const uint64_t dims[3] = {(uint64_t)KP, (uint64_t)d.L, (uint64_t)d.BN};
    const uint64_t strides[2] = {(uint64_t)KP * 2, (uint64_t)d.L * KP * 2};
    const uint32_t box[3] = {64, FA_BM, 1};
    AACONV_TRY(make_tmap_bf16(&tq, qa, 3, dims, strides, box, nullptr));
    AACONV_TRY(make_tmap_bf16(&tk, ka, 3, dims, strides, box, nullptr));
    const uint64_t vdims[2] = {(uint64_t)Lp, (uint64_t)d.BN * FA_DV};
    const uint64_t vstr[1] = {(uint64_t)Lp * 2};
    const uint32_t vbox[2] = {64, FA_DV};
    AACONV_TRY(make_tmap_bf16(&tv, vt, 2, vdims, vstr, vbox, nullptr));
  }
  const int ksteps = cdiv(kdim, 16);
  switch (KP / 64) {
    case 1: return launch_fwd<1>(d, tq, tk, tv, o, lse, ksteps, st);
    case 2: return launch_fwd<2>(d, tq, tk, tv, o, lse, ksteps, st);
    default: return launch_fwd<3>(d, tq, tk, tv, o, lse, ksteps, st);
  }
}

}  // namespace aaconv
