// bf16 tcgen05 attention core of AAConv2d -- backward (recompute), and the shared operand builder.
//
// Augmented operands (one row per query / key position, bf16, KP columns, K-major):
//   Qa[row] = [ c*q | c*Aq | c*Bq | -lse2_hi | -lse2_lo | 0.. | dO | -delta_hi | -delta_lo | 0.. ]
//   Ka[row] = [  k  | 1hot(x') | 1hot(y') |   1   |    1    | 0.. |  v |     1    |     1    | 0.. ]
//              `------------- S block: columns [0, C1) ------------'  `-- V block: [C1, C1+16) --'
//   c = log2(e);  Aq[x'] = q.key_rel_w[:, x'-x+W-1];  Bq[y'] = q.key_rel_h[:, y'-y+H-1]   (rel_to_abs as an
//   index computation, attn_aug_conv.py:43-63);  lse2 = c*lse split into two bf16 so the sum is exact to ~1e-4.
// so that two MMAs give, with no per-element subtraction on the CUDA cores,
//   S'[q,k]  = Qa[q,0:C1].Ka[k,0:C1]      = log2e * (logit[q,k] - lse[q])      ->  P = 2^S'
//   dP'[q,k] = Qa[q,C1:C1+16].Ka[k,C1:..] = dO[q].v[k] - delta[q]              ->  dS = P * dP'
// and the same shared-memory tiles, viewed MN-major, are the B operands of
//   dV += P^T dO,  dK += dS^T Qa[:, 0:dkh]  (key-stationary kernel),  dQa += dS Ka[:, 0:KD]  (query-stationary kernel).
// dQa[:, dkh:KD] are the gradients of Aq/Bq, i.e. the relative-logit gradients summed over key rows / columns.
//
// Out-of-range rows are zero-filled by TMA: S' = 0, dP' = 0 -> dS = 0 and P multiplies zero dO rows, so the
// backward needs no masking at all.
#include "tc_common.cuh"
#include "bf16_path.cuh"

namespace aaconv {

using tc::smem_u32;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// ping-pong kernel skeleton: 8 math warps (two warpgroups, one TMEM slot each), 1 TMA warp, 1 MMA warp
// ------------------------------------------------------------------------------------------------
constexpr int PP_THREADS = 352;   // warps 0-7 math, 8 TMA, 9 score-MMA issuer (+TMEM alloc), 10 gradient-MMA issuer
constexpr int PP_BM = 128;        // stationary rows (TMEM lanes)
constexpr int PP_BN = 64;         // streamed rows per tile (TMEM columns per slot)
constexpr int PP_STAGES = 5;      // smem stages of the streamed operand (>= slots + 1)
constexpr int PP_SLOTS = 3;       // TMEM score slots: the MMA warp runs up to two tiles ahead of the math warpgroups
constexpr uint32_t PP_SLOT_COLS = 128;   // per slot: S' [0,64) dP' [64,128);  P aliases S'[0,32), dS aliases dP'[0,32)

template <int KATOMS>
struct __align__(1024) PPSmem {
  bf16 stat[KATOMS][PP_BM * 64];                   // stationary operand, 16 KB per 64-column atom
  bf16 strm[PP_STAGES][KATOMS][PP_BN * 64];        // streamed operand, 8 KB per atom
  uint64_t bar_stat, bar_full[PP_STAGES], bar_empty[PP_STAGES];
  uint64_t bar_s_full[PP_SLOTS], bar_p_ready[PP_SLOTS], bar_slot_free[PP_SLOTS], bar_final, bar_a_ready;
  uint32_t tmem_base;
};

__host__ __device__ constexpr uint32_t idesc_bmn(int M, int N) { return tc::idesc_bf16_f32(M, N) | (1u << 16); }

// Shared skeleton of the two kernels.
//   warp 8: TMA producer;  warp 9: score-MMA issuer;  warp 10: gradient-MMA issuer;  warps 0-7: two math warpgroups,
//   WG w takes tiles j = w (mod 2).  Tile j uses TMEM slot j % nslots.  Two issuing warps, because the tensor pipe idles
//   whenever its single issuer sits in an mbarrier wait (measured with the clock timeline, tools/attn_timeline.py): the
//   score warp runs ahead by up to nslots tiles (it only waits for the K tile and for the slot's previous gradient MMAs,
//   bar_slot_free), the gradient warp follows the math warpgroups (bar_p_ready).
template <int KATOMS>
__device__ __forceinline__ void pp_init(PPSmem<KATOMS>& sm, int warp, int lane, const CUtensorMap* m0, const CUtensorMap* m1) {
  if (threadIdx.x == 0) {
    tc::mbar_init(&sm.bar_stat, 1);
    for (int s = 0; s < PP_STAGES; ++s) { tc::mbar_init(&sm.bar_full[s], 1); tc::mbar_init(&sm.bar_empty[s], 1); }
    for (int s = 0; s < PP_SLOTS; ++s) {
      tc::mbar_init(&sm.bar_s_full[s], 1);
      tc::mbar_init(&sm.bar_p_ready[s], 128);
      tc::mbar_init(&sm.bar_slot_free[s], 1);
    }
    tc::mbar_init(&sm.bar_final, 1);
    tc::mbar_init(&sm.bar_a_ready, 128);
    tc::fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { tc::tma_prefetch_desc(m0); tc::tma_prefetch_desc(m1); }
  if (warp == 9) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
}

template <int KATOMS>
__device__ __forceinline__ void pp_producer(PPSmem<KATOMS>& sm, const CUtensorMap* m_stat, const CUtensorMap* m_strm, int r0,
                                            int bn, int ntiles) {
  tc::mbar_arrive_expect_tx(&sm.bar_stat, KATOMS * PP_BM * 64 * 2);
  for (int a = 0; a < KATOMS; ++a) tc::tma_load_3d(sm.stat[a], m_stat, &sm.bar_stat, a * 64, r0, bn);
  for (int j = 0; j < ntiles; ++j) {
    const int s = j % PP_STAGES, ph = (j / PP_STAGES) & 1;
    tc::mbar_wait(&sm.bar_empty[s], ph ^ 1);
    tc::mbar_arrive_expect_tx(&sm.bar_full[s], KATOMS * PP_BN * 64 * 2);
    for (int a = 0; a < KATOMS; ++a) tc::tma_load_3d(sm.strm[s][a], m_strm, &sm.bar_full[s], a * 64, j * PP_BN, bn);
  }
}

// The stationary operand is the A operand of every score MMA.  A from shared memory costs ~45 cycles per MMA on top of
// N/2 (tools/mma_bench.cu: SS N=64 76 cycles, TS 46), so warpgroup 0 copies the tile once into TMEM (row r -> lane r,
// column c <- elements 2c, 2c+1, the layout P uses) and all score MMAs run in TS mode.
template <int KATOMS>
__device__ __forceinline__ void pp_stationary_to_tmem(PPSmem<KATOMS>& sm, uint32_t tlane_a, int r) {
  tc::mbar_wait(&sm.bar_stat, 0);
#pragma unroll
  for (int a = 0; a < KATOMS; ++a) {
    uint32_t w[32];
    const uint8_t* rowp = reinterpret_cast<const uint8_t*>(sm.stat[a]) + r * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {                        // undo the 128B swizzle: 16-byte chunk c of row r sits at c ^ (r % 8)
      const uint4 v = *reinterpret_cast<const uint4*>(rowp + ((c ^ (r & 7)) << 4));
      w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
    tc::tmem_st_x32(tlane_a + a * 32, w);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  tc::mbar_arrive(&sm.bar_a_ready);
}

// ---- key-stationary: dK, dV ------------------------------------------------------------------------
// TMEM columns: slot s at 128*s: S'^T [0,64) dP'^T [64,128); P^T over [0,32), dS^T over [64,96).  dV at 384, dK at 400.
template <int KATOMS>
__global__ void __launch_bounds__(PP_THREADS, 1) attn_bwd_dkv_tc_kernel(
    const __grid_constant__ CUtensorMap tm_k_stat, const __grid_constant__ CUtensorMap tm_q_strm,
    float* __restrict__ dk, float* __restrict__ dv, bf16* __restrict__ dqkvh, int KPq, int nh, int L, int dkh, int dvh,
    int C1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  PPSmem<KATOMS>& sm = *reinterpret_cast<PPSmem<KATOMS>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = blockIdx.y, k0 = blockIdx.x * PP_BM;
  const int ntiles = (L + PP_BN - 1) / PP_BN;
  const int nks = C1 >> 4;
  // TMEM columns: Ka (bf16) [0, 32*KATOMS); slots of 128 behind it; dV (16) and dK (32) behind the slots
  constexpr int NS = (KATOMS * 32 + 3 * 128 + 48 <= 512) ? 3 : 2;
  constexpr uint32_t COL_SLOT0 = KATOMS * 32, COL_DV = COL_SLOT0 + 128 * NS, COL_DK = COL_DV + 16;

  pp_init(sm, warp, lane, &tm_k_stat, &tm_q_strm);
  const uint32_t tmem = sm.tmem_base;

  if (warp == 8) {
    if (lane == 0) pp_producer(sm, &tm_k_stat, &tm_q_strm, k0, bn, ntiles);
  } else if (warp == 9) {
    constexpr uint32_t idesc_s = tc::idesc_bf16_f32(PP_BM, PP_BN);
    constexpr uint32_t idesc_dv = idesc_bmn(PP_BM, 16);
    constexpr uint32_t idesc_dk = idesc_bmn(PP_BM, 32);
    constexpr uint32_t STAT_ATOM = (PP_BM * 128) >> 4, STRM_ATOM = (PP_BN * 128) >> 4, STAGE = KATOMS * STRM_ATOM;
    const uint32_t stat_lo = tc::desc_lo_k(smem_u32(sm.stat[0]));
    const uint32_t strm_lo = tc::desc_lo_k(smem_u32(sm.strm[0][0]));
    const uint32_t v_lo = tc::desc_lo_mn(smem_u32(sm.strm[0][C1 >> 6]) + (C1 & 63) * 2, PP_BN * 128);
    const uint32_t q_lo = tc::desc_lo_mn(smem_u32(sm.strm[0][0]), PP_BN * 128);
    tc::mbar_wait(&sm.bar_a_ready, 0);
    tc::tc_fence_after();
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % PP_STAGES, slot = j % NS;
      const uint32_t tslot = tmem + COL_SLOT0 + PP_SLOT_COLS * slot;
      tc::mbar_wait(&sm.bar_full[st], (j / PP_STAGES) & 1);
      if (j >= NS) tc::mbar_wait(&sm.bar_slot_free[slot], ((j / NS) - 1) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        tc::issue_ts_ksteps_n<1, STRM_ATOM>(nks, tslot, tslot + 64, tmem, strm_lo + st * STAGE, idesc_s);
        tc::mma_commit(&sm.bar_s_full[slot]);
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    constexpr uint32_t idesc_dv = idesc_bmn(PP_BM, 16);
    constexpr uint32_t idesc_dk = idesc_bmn(PP_BM, 32);
    constexpr uint32_t STRM_ATOM = (PP_BN * 128) >> 4, STAGE = KATOMS * STRM_ATOM;
    const uint32_t v_lo = tc::desc_lo_mn(smem_u32(sm.strm[0][C1 >> 6]) + (C1 & 63) * 2, PP_BN * 128);
    const uint32_t q_lo = tc::desc_lo_mn(smem_u32(sm.strm[0][0]), PP_BN * 128);
    for (int jj = 0; jj < ntiles; ++jj) {
      const int st = jj % PP_STAGES, slot = jj % NS;
      const uint32_t tslot = tmem + COL_SLOT0 + PP_SLOT_COLS * slot;
      tc::mbar_wait(&sm.bar_p_ready[slot], (jj / NS) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t vb = v_lo + st * STAGE, qb = q_lo + st * STAGE;
#pragma unroll
        for (int ks = 0; ks < PP_BN / 16; ++ks) {
          tc::mma_ts(tmem + COL_DV, tslot + ks * 8, tc::desc64(vb + ks * 128), idesc_dv, (jj > 0 || ks > 0) ? 1u : 0u);
          tc::mma_ts(tmem + COL_DK, tslot + 64 + ks * 8, tc::desc64(qb + ks * 128), idesc_dk, (jj > 0 || ks > 0) ? 1u : 0u);
        }
        tc::mma_commit(&sm.bar_empty[st]);
        tc::mma_commit(&sm.bar_slot_free[slot]);
        if (jj == ntiles - 1) tc::mma_commit(&sm.bar_final);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ===================== math warpgroups: thread == key row == TMEM lane =====================
    const int wg = warp >> 2;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    if (wg == 0) pp_stationary_to_tmem(sm, tlane, (warp & 3) * 32 + lane);
    uint32_t rs[32], rd[32], pp[16], pd[16];
    for (int j = wg; j < ntiles; j += 2) {
      const int slot = j % NS;
      const uint32_t tslot = tlane + COL_SLOT0 + PP_SLOT_COLS * slot;
      tc::mbar_wait(&sm.bar_s_full[slot], (j / NS) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int c = 0; c < PP_BN / 32; ++c) {
        tc::tmem_ld_x32(tslot + c * 32, rs);
        tc::tmem_ld_x32(tslot + 64 + c * 32, rd);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = tc::ex2f(__uint_as_float(rs[2 * i])), p1 = tc::ex2f(__uint_as_float(rs[2 * i + 1]));
          pp[i] = tc::pack_bf16x2(p0, p1);
          pd[i] = tc::pack_bf16x2(p0 * __uint_as_float(rd[2 * i]), p1 * __uint_as_float(rd[2 * i + 1]));
        }
        tc::tmem_st_x16(tslot + c * 16, pp);          // P^T over S'^T[0,32): those columns are already in registers
        tc::tmem_st_x16(tslot + 64 + c * 16, pd);     // dS^T over dP'^T[0,32)
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&sm.bar_p_ready[slot]);
    }
    // epilogue: after the last gradient MMAs WG0 writes dK, WG1 writes dV
    tc::mbar_wait(&sm.bar_final, 0);
    tc::tc_fence_after();
    const int kj = k0 + (warp & 3) * 32 + lane;
    const size_t row = (size_t)bn * L + kj;
    const float LN2 = 0.6931471805599453f;               // Qa carries log2(e)*q
    // packed bf16 destination: pixel (b, kj), columns [nh*dkh + n*dkh, ..) for dk and [2*nh*dkh + n*dvh, ..) for dv
    const int b = bn / nh, n = bn - b * nh;
    bf16* prow = dqkvh ? dqkvh + ((size_t)b * L + kj) * KPq : nullptr;
    if (wg == 0) {
      tc::tmem_ld_x32(tlane + COL_DK, rs);
      tc::tmem_ld_wait();
      if (kj < L) {
        if (prow) {
          bf16* dst = prow + nh * dkh + n * dkh;
          if (((dkh | KPq) & 3) == 0) {                    // 8-byte aligned groups of four
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              if (e < dkh) {
                uint2 w;
                w.x = tc::pack_bf16x2(__uint_as_float(rs[e]) * LN2, __uint_as_float(rs[e + 1]) * LN2);
                w.y = tc::pack_bf16x2(__uint_as_float(rs[e + 2]) * LN2, __uint_as_float(rs[e + 3]) * LN2);
                *reinterpret_cast<uint2*>(dst + e) = w;
              }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e < dkh) dst[e] = __float2bfloat16(__uint_as_float(rs[e]) * LN2);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e < dkh) dk[row * dkh + e] = __uint_as_float(rs[e]) * LN2;
        }
      }
    } else {
      tc::tmem_ld_x16(tlane + COL_DV, pp);
      tc::tmem_ld_wait();
      if (kj < L) {
        if (prow) {
          bf16* dst = prow + 2 * nh * dkh + n * dvh;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (e < dvh) dst[e] = __float2bfloat16(__uint_as_float(pp[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (e < dvh) dv[row * dvh + e] = __uint_as_float(pp[e]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc<512>(tmem);
}

// ---- query-stationary: dQa (content gradient + gradients of the relative rows Aq, Bq) ---------------
// TMEM columns: slot s at 128*s: S' [0,64) dP' [64,128); dS over [64,96).  dQa (NQ <= 160 columns) behind the slots:
// three slots when NQ <= 128 (dQa at 384), two otherwise (dQa at 256).
template <int KATOMS>
__global__ void __launch_bounds__(PP_THREADS, 1) attn_bwd_dq_tc_kernel(
    const __grid_constant__ CUtensorMap tm_q_stat, const __grid_constant__ CUtensorMap tm_k_strm,
    float* __restrict__ dqa, int L, int KD, int NQ, int C1, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  PPSmem<KATOMS>& sm = *reinterpret_cast<PPSmem<KATOMS>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = blockIdx.y, q0 = blockIdx.x * PP_BM;
  const int ntiles = (L + PP_BN - 1) / PP_BN;
  const int nks = C1 >> 4;
  // TMEM columns: Qa (bf16) [0, 32*KATOMS); slots of 128 behind it; dQa (NQ <= 160) behind the slots
  const int NS = (KATOMS * 32 + 3 * 128 + NQ <= 512) ? 3 : 2;
  const uint32_t COL_SLOT0 = KATOMS * 32, COL_DQ = COL_SLOT0 + PP_SLOT_COLS * NS;

  pp_init(sm, warp, lane, &tm_q_stat, &tm_k_strm);
  const uint32_t tmem = sm.tmem_base;
  // optional timeline of CTA (0,0): dbg[event][tile], clock64 stamps (tools/attn_timeline.py)
  const bool rec = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
#define PP_STAMP(ev, tile) do { if (rec) dbg[(ev) * 64 + (tile)] = clock64(); } while (0)   // 12 events x 64 tiles

  if (warp == 8) {
    if (lane == 0) pp_producer(sm, &tm_q_stat, &tm_k_strm, q0, bn, ntiles);
  } else if (warp == 9) {
    constexpr uint32_t idesc_s = tc::idesc_bf16_f32(PP_BM, PP_BN);
    const uint32_t idesc_dq = idesc_bmn(PP_BM, NQ);
    constexpr uint32_t STAT_ATOM = (PP_BM * 128) >> 4, STRM_ATOM = (PP_BN * 128) >> 4, STAGE = KATOMS * STRM_ATOM;
    const uint32_t stat_lo = tc::desc_lo_k(smem_u32(sm.stat[0]));
    const uint32_t strm_lo = tc::desc_lo_k(smem_u32(sm.strm[0][0]));
    const uint32_t k_lo = tc::desc_lo_mn(smem_u32(sm.strm[0][0]), PP_BN * 128);
    tc::mbar_wait(&sm.bar_a_ready, 0);
    tc::tc_fence_after();
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % PP_STAGES, slot = j % NS;
      const uint32_t tslot = tmem + COL_SLOT0 + PP_SLOT_COLS * slot;
      PP_STAMP(8, j);
      tc::mbar_wait(&sm.bar_full[st], (j / PP_STAGES) & 1);
      if (j >= NS) tc::mbar_wait(&sm.bar_slot_free[slot], ((j / NS) - 1) & 1);
      tc::tc_fence_after();
      PP_STAMP(0, j);
      if (tc::elect_one()) {
        tc::issue_ts_ksteps_n<1, STRM_ATOM>(nks, tslot, tslot + 64, tmem, strm_lo + st * STAGE, idesc_s);
        tc::mma_commit(&sm.bar_s_full[slot]);
      }
      __syncwarp();
      PP_STAMP(1, j);
    }
  } else if (warp == 10) {
    const uint32_t idesc_dq = idesc_bmn(PP_BM, NQ);
    constexpr uint32_t STRM_ATOM = (PP_BN * 128) >> 4, STAGE = KATOMS * STRM_ATOM;
    const uint32_t k_lo = tc::desc_lo_mn(smem_u32(sm.strm[0][0]), PP_BN * 128);
    for (int jj = 0; jj < ntiles; ++jj) {
      const int st = jj % PP_STAGES, slot = jj % NS;
      const uint32_t tslot = tmem + COL_SLOT0 + PP_SLOT_COLS * slot;
      tc::mbar_wait(&sm.bar_p_ready[slot], (jj / NS) & 1);
      tc::tc_fence_after();
      PP_STAMP(2, jj);
      if (tc::elect_one()) {
        const uint32_t kb = k_lo + st * STAGE;
#pragma unroll
        for (int ks = 0; ks < PP_BN / 16; ++ks)
          tc::mma_ts(tmem + COL_DQ, tslot + 64 + ks * 8, tc::desc64(kb + ks * 128), idesc_dq, (jj > 0 || ks > 0) ? 1u : 0u);
        tc::mma_commit(&sm.bar_empty[st]);
        tc::mma_commit(&sm.bar_slot_free[slot]);
        if (jj == ntiles - 1) tc::mma_commit(&sm.bar_final);
      }
      __syncwarp();
      PP_STAMP(9, jj);
    }
  } else if (warp < 8) {
    const int wg = warp >> 2;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    if (wg == 0) pp_stationary_to_tmem(sm, tlane, (warp & 3) * 32 + lane);
    uint32_t rs[32], rd[32], pd[16];
    for (int j = wg; j < ntiles; j += 2) {
      const int slot = j % NS;
      const uint32_t tslot = tlane + COL_SLOT0 + PP_SLOT_COLS * slot;
      if ((warp & 3) == 0) PP_STAMP(3, j);
      tc::mbar_wait(&sm.bar_s_full[slot], (j / NS) & 1);
      tc::tc_fence_after();
      if ((warp & 3) == 0) PP_STAMP(4, j);
#pragma unroll
      for (int c = 0; c < PP_BN / 32; ++c) {
        tc::tmem_ld_x32(tslot + c * 32, rs);
        tc::tmem_ld_x32(tslot + 64 + c * 32, rd);
        tc::tmem_ld_wait();
        if ((warp & 3) == 0 && c == 0) PP_STAMP(5, j);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = tc::ex2f(__uint_as_float(rs[2 * i])), p1 = tc::ex2f(__uint_as_float(rs[2 * i + 1]));
          pd[i] = tc::pack_bf16x2(p0 * __uint_as_float(rd[2 * i]), p1 * __uint_as_float(rd[2 * i + 1]));
        }
        tc::tmem_st_x16(tslot + 64 + c * 16, pd);     // dS over dP'[0,32): already in registers
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&sm.bar_p_ready[slot]);
      if ((warp & 3) == 0) PP_STAMP(6, j);
    }
    if ((warp & 3) == 0) PP_STAMP(7, wg);
    tc::mbar_wait(&sm.bar_final, 0);
    if ((warp & 3) == 0) PP_STAMP(7, 2 + wg);
    tc::tc_fence_after();
    // dQa rows -> global (B,nh,L,KD) fp32; the two warpgroups split the columns in 32-wide chunks
    const int qi = q0 + (warp & 3) * 32 + lane;
    const size_t row = (size_t)bn * L + qi;
    for (int c0 = wg * 32; c0 < NQ; c0 += 64) {
      tc::tmem_ld_x32(tlane + COL_DQ + c0, rs);
      tc::tmem_ld_wait();
      if (qi < L) {
        if ((KD & 3) == 0) {                         // 16-byte stores (row stride and column offset are multiples of 4 floats)
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (c0 + e < KD)
              *reinterpret_cast<float4*>(dqa + row * KD + c0 + e) =
                  make_float4(__uint_as_float(rs[e]), __uint_as_float(rs[e + 1]), __uint_as_float(rs[e + 2]), __uint_as_float(rs[e + 3]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e < KD) dqa[row * KD + c0 + e] = __uint_as_float(rs[e]);
        }
      }
    }
  }
  if ((warp & 3) == 0 && warp < 8) PP_STAMP(7, 4 + (warp >> 2));
#undef PP_STAMP
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc<512>(tmem);
}

// debug hook (tools/attn_timeline.py): device buffer of 8 x 64 clock stamps written by CTA (0,0) of the dq kernel
long long* g_attn_dbg = nullptr;
extern "C" void aaconv_debug_set_timeline(void* p) { g_attn_dbg = static_cast<long long*>(p); }

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int make_aug_maps(const Dims& d, const void* t, int KP, uint32_t box_rows, CUtensorMap* out) {
  const uint64_t dims[3] = {(uint64_t)KP, (uint64_t)d.L, (uint64_t)d.BN};
  const uint64_t strides[2] = {(uint64_t)KP * 2, (uint64_t)d.L * KP * 2};
  const uint32_t box[3] = {64, box_rows, 1};
  return make_tmap_bf16(out, t, 3, dims, strides, box, nullptr);
}

template <int KATOMS>
static int launch_bwd(const Dims& d, const AugLayout& a, const void* qa, const void* ka, float* dqa, float* dk,
                      float* dv, void* dqkvh, int KPq, cudaStream_t st) {
  CUtensorMap tq_stat, tk_strm, tk_stat, tq_strm;
  AACONV_TRY(make_aug_maps(d, qa, a.KP, PP_BM, &tq_stat));
  AACONV_TRY(make_aug_maps(d, ka, a.KP, PP_BN, &tk_strm));
  AACONV_TRY(make_aug_maps(d, ka, a.KP, PP_BM, &tk_stat));
  AACONV_TRY(make_aug_maps(d, qa, a.KP, PP_BN, &tq_strm));
  const size_t smem = sizeof(PPSmem<KATOMS>) + 1024;
  dim3 grid(cdiv(d.L, PP_BM), d.BN);
  {
    auto kern = attn_bwd_dkv_tc_kernel<KATOMS>;
    AACONV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, PP_THREADS, smem, AACONV_ST(st)>>>(tk_stat, tq_strm, dk, dv, static_cast<bf16*>(dqkvh), KPq, d.nh, d.L, d.dkh, d.dvh, a.C1);
    AACONV_LAUNCH_OK("attn_bwd_dkv_tc");
  }
  {
    auto kern = attn_bwd_dq_tc_kernel<KATOMS>;
    AACONV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, PP_THREADS, smem, AACONV_ST(st)>>>(tq_stat, tk_strm, dqa, d.L, a.KD, a.NQ, a.C1, g_attn_dbg);
    AACONV_LAUNCH_OK("attn_bwd_dq_tc");
  }
  return 0;
}

// qa/ka: mode-1 augmented operands.  dqa (B,nh,L,KD) fp32; dk (B,nh,L,dkh); dv (B,nh,L,dvh).
int tc_attn_bwd(const Dims& d, const void* qa, const void* ka, float* dqa, float* dk, float* dv, void* dqkvh, int KPq,
                cudaStream_t st) {
  AACONV_TRY(aug_supported(d));
  const AugLayout a = aug_layout(d);
  switch (a.KP / 64) {
    case 1: return launch_bwd<1>(d, a, qa, ka, dqa, dk, dv, dqkvh, KPq, st);
    case 2: return launch_bwd<2>(d, a, qa, ka, dqa, dk, dv, dqkvh, KPq, st);
    default: return launch_bwd<3>(d, a, qa, ka, dqa, dk, dv, dqkvh, KPq, st);
  }
}

// ------------------------------------------------------------------------------------------------
// post-processing of dQa: total dq, and the abs->rel scatter for the key_rel gradients
// ------------------------------------------------------------------------------------------------
// dq[row, e] = dQa[row, e] + sum_x' krw[e, x'-x+W-1] dQa[row, dkh+x'] + sum_y' krh[e, y'-y+H-1] dQa[row, dkh+W+y']
__global__ void aug_bwd_dq_kernel(const float* __restrict__ dqa, const float* __restrict__ krw,
                                  const float* __restrict__ krh, float* __restrict__ dq, size_t rows, int L, int H,
                                  int W, int dkh, int KD, int relative) {
  extern __shared__ float sm[];
  const int RW = 2 * W - 1, RH = 2 * H - 1;
  float* kw = sm;
  float* kh = sm + (relative ? dkh * RW : 0);
  if (relative) {
    for (int i = threadIdx.x; i < dkh * RW; i += blockDim.x) kw[i] = krw[i];
    for (int i = threadIdx.x; i < dkh * RH; i += blockDim.x) kh[i] = krh[i];
  }
  __syncthreads();
  const size_t total = rows * dkh;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / dkh;
    const int e = (int)(i - row * dkh);
    const float* g = dqa + row * KD;
    float s = __ldg(g + e);
    if (relative) {
      const int l = (int)(row % L), y = l / W, x = l - y * W;
      const float* tw = kw + e * RW + (W - 1 - x);
      const float* th = kh + e * RH + (H - 1 - y);
      for (int xp = 0; xp < W; ++xp) s = fmaf(tw[xp], __ldg(g + dkh + xp), s);
      for (int yp = 0; yp < H; ++yp) s = fmaf(th[yp], __ldg(g + dkh + W + yp), s);
    }
    dq[i] = s;
  }
}

int aug_bwd_dq(const Dims& d, const float* dqa, const float* krw, const float* krh, float* dq, cudaStream_t st) {
  const AugLayout a = aug_layout(d);
  const size_t rows = (size_t)d.BN * d.L;
  const size_t smem = d.relative ? (size_t)d.dkh * (d.RW + d.RH) * sizeof(float) : 0;
  AACONV_CUDA_OK(cudaFuncSetAttribute(aug_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<size_t>((rows * d.dkh + 255) / 256, 148 * 16);
  aug_bwd_dq_kernel<<<grid, 256, smem, AACONV_ST(st)>>>(dqa, krw, krh, dq, rows, d.L, d.H, d.W, d.dkh, a.KD, d.relative);
  AACONV_LAUNCH_OK("aug_bwd_dq");
  return 0;
}

}  // namespace aaconv
