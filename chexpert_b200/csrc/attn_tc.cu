// bf16 tcgen05 attention core of AAConv2d (forward): TMA-fed Qa.Ka^T into TMEM, online softmax from TMEM,
// P.[v|1] from TMEM-resident P.  Reference rows a3-a8 (attn_aug_conv.py:75-91).
//
// Operands are the augmented Qa / Ka rows built by aug_build_fwd (layout: attn_tc_bwd.cu header):
//   S'[q,k] = Qa[q, 0:C1] . Ka[k, 0:C1] = log2(e) * (q.k + Rw[q, x'-x+W-1] + Rh[q, y'-y+H-1])
// (the backward-only columns of Qa are still zero), so the relative logits are part of the QK^T contraction and
// nothing of size (HW x HW) ever leaves TMEM.  The V block of the same Ka tile, viewed MN-major, is the B operand
// of P.V; its "ones" column makes that MMA produce the softmax row sum as well.
//
// Structure: one CTA = 128 queries of one (batch, head); two softmax warpgroups ping-pong over the key tiles
// (WG0 even tiles, WG1 odd tiles), each with its own TMEM slot (S, P, O) and its own running maximum; the MMA warp
// issues S of tile j+1 while tile j is in the softmax, so the tensor pipe, the MUFU pipe and the TMEM reads of
// the two warpgroups overlap.  The two partial (m, l, O) are merged at the end.  The maximum is only raised when
// it grows by more than 2^8 (lazy rescale), which removes almost all O read-modify-write round trips.
#include "tc_common.cuh"
#include "bf16_path.cuh"

namespace aaconv {

using tc::smem_u32;
typedef __nv_bfloat16 bf16;

constexpr int FA_BM = 128;       // queries per CTA (TMEM lanes)
constexpr int FA_BN = 128;       // keys per tile
template <int KATOMS> struct FaStages { static constexpr int value = KATOMS >= 3 ? 3 : 4; };   // K smem stages (>= slots; 48 KB each at KATOMS = 3)
constexpr int FA_SLOTS = 3;      // TMEM score slots: the MMA warp runs up to two tiles ahead of the softmax warpgroups
constexpr int FA_THREADS = 352;  // warps 0-3 softmax WG0, 4-7 softmax WG1, warp 8 TMA, warp 9 score MMAs (+TMEM alloc), warp 10 P.V MMAs
constexpr uint32_t FA_SLOT = 128;   // TMEM: Qa (bf16) [0, 32*KATOMS); slot s: S at 32*KATOMS + 128*s (P, bf16, aliases its first 64 columns);
                                    // O of warpgroup w behind the slots at + 16*w
constexpr float FA_RESCALE_THRESHOLD = 8.f;                          // log2 units

template <int KATOMS>
struct __align__(1024) FwdSmem {
  static constexpr int FA_STAGES = FaStages<KATOMS>::value;
  bf16 q[KATOMS][FA_BM * 64];                 // K-major, 128 B rows, 128B swizzle (one atom = 64 k-elements)
  bf16 k[FA_STAGES][KATOMS][FA_BN * 64];
  float xch[FA_BM][18];                       // WG1 -> WG0 hand-over of (m, O[0..16))
  uint64_t bar_q, bar_a_ready, bar_full[FA_STAGES], bar_empty[FA_STAGES], bar_s_full[FA_SLOTS], bar_p_ready[FA_SLOTS], bar_slot_free[FA_SLOTS], bar_o_done[2];
  uint32_t tmem_base;
};

template <int KATOMS>
__global__ void __launch_bounds__(FA_THREADS, 1) attn_fwd_tc_kernel(
    const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, float* __restrict__ o,
    float* __restrict__ lse, int L, int dvh, int C1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FwdSmem<KATOMS>& sm = *reinterpret_cast<FwdSmem<KATOMS>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = blockIdx.y, q0 = blockIdx.x * FA_BM;
  const int ntiles = (L + FA_BN - 1) / FA_BN;
  constexpr int FA_STAGES = FaStages<KATOMS>::value;

  if (threadIdx.x == 0) {
    tc::mbar_init(&sm.bar_q, 1);
    tc::mbar_init(&sm.bar_a_ready, 128);
    for (int s = 0; s < FA_STAGES; ++s) { tc::mbar_init(&sm.bar_full[s], 1); tc::mbar_init(&sm.bar_empty[s], 1); }
    for (int s = 0; s < FA_SLOTS; ++s) {
      tc::mbar_init(&sm.bar_s_full[s], 1);
      tc::mbar_init(&sm.bar_p_ready[s], 128);
      tc::mbar_init(&sm.bar_slot_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) tc::mbar_init(&sm.bar_o_done[s], 1);
    tc::fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); }
  if (warp == 9) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  constexpr uint32_t COL_SLOT0 = KATOMS * 32, FA_COL_O = COL_SLOT0 + FA_SLOT * FA_SLOTS;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&sm.bar_q, KATOMS * FA_BM * 64 * 2);
      for (int a = 0; a < KATOMS; ++a) tc::tma_load_3d(sm.q[a], &tm_q, &sm.bar_q, a * 64, q0, bn);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % FA_STAGES, ph = (j / FA_STAGES) & 1;
        tc::mbar_wait(&sm.bar_empty[s], ph ^ 1);
        tc::mbar_arrive_expect_tx(&sm.bar_full[s], KATOMS * FA_BN * 64 * 2);
        for (int a = 0; a < KATOMS; ++a) tc::tma_load_3d(sm.k[s][a], &tm_k, &sm.bar_full[s], a * 64, j * FA_BN, bn);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (whole warp runs the uniform loop, one elected lane issues) =====================
    constexpr uint32_t idesc_s = tc::idesc_bf16_f32(FA_BM, FA_BN);
    constexpr uint32_t idesc_o = tc::idesc_bf16_f32(FA_BM, 16) | (1u << 16);      // B (the V block) is MN-major
    constexpr uint32_t K_ATOM = (FA_BN * 128) >> 4, K_STAGE = KATOMS * K_ATOM;
    const uint32_t k_lo = tc::desc_lo_k(smem_u32(sm.k[0][0]));
    const uint32_t v_lo = tc::desc_lo_mn(smem_u32(sm.k[0][C1 >> 6]) + (C1 & 63) * 2, FA_BN * 128);
    const int nks = C1 >> 4;
    tc::mbar_wait(&sm.bar_a_ready, 0);        // Qa copied to TMEM by warpgroup 0: score MMAs run in TS mode (tools/mma_bench.cu)
    tc::tc_fence_after();
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % FA_STAGES, slot = j % FA_SLOTS;
      tc::mbar_wait(&sm.bar_full[st], (j / FA_STAGES) & 1);
      if (j >= FA_SLOTS) tc::mbar_wait(&sm.bar_slot_free[slot], ((j / FA_SLOTS) - 1) & 1);   // P.V of the slot's previous tile done
      tc::tc_fence_after();
      if (tc::elect_one()) {
        tc::issue_ts_ksteps_n<0, K_ATOM>(nks, tmem + COL_SLOT0 + FA_SLOT * slot, 0u, tmem, k_lo + st * K_STAGE, idesc_s);
        tc::mma_commit(&sm.bar_s_full[slot]);
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // ===================== P.V issuer: a second issuing warp keeps the tensor pipe fed while the other one waits =====================
    constexpr uint32_t idesc_o = tc::idesc_bf16_f32(FA_BM, 16) | (1u << 16);      // B (the V block) is MN-major
    constexpr uint32_t K_ATOM = (FA_BN * 128) >> 4, K_STAGE = KATOMS * K_ATOM;
    const uint32_t v_lo = tc::desc_lo_mn(smem_u32(sm.k[0][C1 >> 6]) + (C1 & 63) * 2, FA_BN * 128);
    for (int jj = 0; jj < ntiles; ++jj) {
      const int st = jj % FA_STAGES, slot = jj % FA_SLOTS, wgo = jj & 1;
      tc::mbar_wait(&sm.bar_p_ready[slot], (jj / FA_SLOTS) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t vb = v_lo + st * K_STAGE, tp = tmem + COL_SLOT0 + FA_SLOT * slot, to = tmem + FA_COL_O + 16 * wgo;
#pragma unroll
        for (int ks = 0; ks < FA_BN / 16; ++ks)    // 16 keys = 16 rows of 128 B = 2048 B = 128 descriptor units
          tc::mma_ts(to, tp + ks * 8, tc::desc64(vb + ks * 128), idesc_o, (jj > 1 || ks > 0) ? 1u : 0u);
        tc::mma_commit(&sm.bar_empty[st]);
        tc::mma_commit(&sm.bar_slot_free[slot]);
        tc::mma_commit(&sm.bar_o_done[wgo]);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ===================== softmax warpgroups (thread == query row == TMEM lane) =====================
    const int wg = warp >> 2;
    const int rowi = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tout = tlane + FA_COL_O + 16 * wg;
    if (wg == 0) {                             // stationary Qa tile: shared memory -> TMEM (row -> lane, column c <- elements 2c, 2c+1)
      tc::mbar_wait(&sm.bar_q, 0);
#pragma unroll
      for (int a = 0; a < KATOMS; ++a) {
        uint32_t w[32];
        const uint8_t* rowp = reinterpret_cast<const uint8_t*>(sm.q[a]) + rowi * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {          // undo the 128B swizzle: 16-byte chunk c of row r sits at c ^ (r % 8)
          const uint4 v = *reinterpret_cast<const uint4*>(rowp + ((c ^ (rowi & 7)) << 4));
          w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
        tc::tmem_st_x32(tlane + a * 32, w);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&sm.bar_a_ready);
    }
    float m = -INFINITY;                      // reference maximum of this warpgroup's exponentials (log2 units)
    uint32_t r[4][32];
    int nmine = 0;
    for (int j = wg; j < ntiles; j += 2, ++nmine) {
      const int it = j >> 1, slot = j % FA_SLOTS;
      const uint32_t tslot = tlane + COL_SLOT0 + FA_SLOT * slot;
      tc::mbar_wait(&sm.bar_s_full[slot], (j / FA_SLOTS) & 1);
      tc::tc_fence_after();
      tc::tmem_ld_x32(tslot + 0, r[0]);
      tc::tmem_ld_x32(tslot + 32, r[1]);
      tc::tmem_ld_wait();
      tc::tmem_ld_x32(tslot + 64, r[2]);
      tc::tmem_ld_x32(tslot + 96, r[3]);
      const int nvalid = L - j * FA_BN;       // >= FA_BN for every tile but (possibly) the last
      if (nvalid < FA_BN) {                   // zero-filled keys past L must not count: logit -> -inf
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= nvalid) r[c][i] = 0xff800000u;
      }
      float mt = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) mt = fmaxf(mt, fmaxf(__uint_as_float(r[c][i]), __uint_as_float(r[c][i + 1])));
      tc::tmem_ld_wait();
      if (nvalid < FA_BN) {
#pragma unroll
        for (int c = 2; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= nvalid) r[c][i] = 0xff800000u;
      }
#pragma unroll
      for (int c = 2; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) mt = fmaxf(mt, fmaxf(__uint_as_float(r[c][i]), __uint_as_float(r[c][i + 1])));
      if (it > 0) {
        tc::mbar_wait(&sm.bar_o_done[wg], (it - 1) & 1);      // P.V of this slot's previous tile has consumed P, updated O
        tc::tc_fence_after();
        if (__any_sync(0xffffffffu, mt > m + FA_RESCALE_THRESHOLD)) {
          const float m_new = fmaxf(m, mt);
          const float alpha = tc::ex2f(m - m_new);
          uint32_t ov[16];
          tc::tmem_ld_x16(tout, ov);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tc::tmem_st_x16(tout, ov);
          m = m_new;
        }
      } else {
        m = mt;
      }
      const float mneg = -m;
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        uint32_t pk[32];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = tc::ex2f(__uint_as_float(r[c + h][2 * i]) + mneg);
            const float p1 = tc::ex2f(__uint_as_float(r[c + h][2 * i + 1]) + mneg);
            pk[h * 16 + i] = tc::pack_bf16x2(p0, p1);
          }
        tc::tmem_st_x32(tslot + c * 16, pk);          // P over S[0,64): the whole S row is already in registers
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&sm.bar_p_ready[slot]);
    }
    // ---- merge the two warpgroups' partial results ----
    uint32_t ov[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ov[i] = 0u;
    if (nmine > 0) {
      tc::mbar_wait(&sm.bar_o_done[wg], (nmine - 1) & 1);
      tc::tc_fence_after();
      tc::tmem_ld_x16(tout, ov);
      tc::tmem_ld_wait();
    }
    if (wg == 1) {
      sm.xch[rowi][0] = m;
#pragma unroll
      for (int i = 0; i < 16; ++i) sm.xch[rowi][1 + i] = __uint_as_float(ov[i]);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (wg == 0) {
      const float m1 = sm.xch[rowi][0];
      const float mm = fmaxf(m, m1);
      const float a0 = tc::ex2f(m - mm), a1 = (m1 == -INFINITY) ? 0.f : tc::ex2f(m1 - mm);
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = a0 * __uint_as_float(ov[i]) + a1 * sm.xch[rowi][1 + i];
      const int qi = q0 + rowi;
      if (qi < L) {
        float l = 1.f;                                   // the ones column of the V block: softmax row sum
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (e == dvh) l = acc[e];
        const float inv = 1.f / l;
        const size_t row = (size_t)bn * L + qi;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (e < dvh) o[row * dvh + e] = acc[e] * inv;
        lse[row] = (mm + log2f(l)) * 0.6931471805599453f;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc<512>(tmem);
}

int tc_attn_supported(const Dims& d) { return aug_supported(d); }

template <int KATOMS>
static int launch_fwd(const Dims& d, const AugLayout& a, const CUtensorMap& tq, const CUtensorMap& tk, float* o, float* lse,
                      cudaStream_t st) {
  const size_t smem = sizeof(FwdSmem<KATOMS>) + 1024;
  auto kern = attn_fwd_tc_kernel<KATOMS>;
  AACONV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(d.L, FA_BM), d.BN);
  kern<<<grid, FA_THREADS, smem, AACONV_ST(st)>>>(tq, tk, o, lse, d.L, d.dvh, a.C1);
  AACONV_LAUNCH_OK("attn_fwd_tc");
  return 0;
}

// qa, ka: augmented operands of aug_build_fwd (backward columns of qa still zero).  o (B,nh,L,dvh), lse (B,nh,L) fp32.
int tc_attn_fwd(const Dims& d, const void* qa, const void* ka, float* o, float* lse, cudaStream_t st) {
  AACONV_TRY(aug_supported(d));
  const AugLayout a = aug_layout(d);
  CUtensorMap tq, tk;
  const uint64_t dims[3] = {(uint64_t)a.KP, (uint64_t)d.L, (uint64_t)d.BN};
  const uint64_t strides[2] = {(uint64_t)a.KP * 2, (uint64_t)d.L * a.KP * 2};
  const uint32_t box[3] = {64, FA_BM, 1};
  AACONV_TRY(make_tmap_bf16(&tq, qa, 3, dims, strides, box, nullptr));
  AACONV_TRY(make_tmap_bf16(&tk, ka, 3, dims, strides, box, nullptr));
  switch (a.KP / 64) {
    case 1: return launch_fwd<1>(d, a, tq, tk, o, lse, st);
    case 2: return launch_fwd<2>(d, a, tq, tk, o, lse, st);
    default: return launch_fwd<3>(d, a, tq, tk, o, lse, st);
  }
}

}  // namespace aaconv
