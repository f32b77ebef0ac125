// bf16 tcgen05 implicit GEMMs for the dense contractions of AAConv2d:
//   fprop : y[:, :Cc] = conv3x3_s2(x)  and  qkv = 1x1_s2(x)   in ONE kernel (the 1x1 projection is the centre
//           tap of the 3x3 stencil, attn_aug_conv.py:34-35, so both read the same TMA boxes of x)
//   dgrad : dx = conv^T(dy[:, :Cc]) + scatter(Wqkv^T dqkv), one launch per input-pixel residue class
// Activations are packed once to channels-last bf16 (NHWC); a tile of M = r x Wt output pixels is one TMA box
// [64 ch, Wt, r, 1] per (tap, 64-channel atom), landing K-major / 128B-swizzled exactly as tcgen05 wants it.
// Strided taps are addressed through per-residue tensor maps (dims = pixels of one (h%s, w%s) class) so that
// every box has unit element strides and image borders are TMA zero-fill.
#include <algorithm>
#include <cstring>
#include <vector>
#include "tc_common.cuh"
#include "bf16_path.cuh"

namespace aaconv {

using tc::smem_u32;
typedef __nv_bfloat16 bf16;

constexpr int PG_THREADS = 192;     // warps 0-3 epilogue, warp 4 TMA, warp 5 MMA
// TWO CTAs per SM: 3 stages (96 KB) and 2 accumulator chunks (256 TMEM columns) each.  One CTA per SM with 6 stages and 4
// chunks left the second of 1.5 (fprop) and the fourth of 3.03 (dgrad) waves almost empty and nothing to run under a CTA's
// epilogue; with half-size CTAs the tail is half as long and one CTA's epilogue overlaps the other's main loop.
// When the whole launch fits one CTA per SM (Transition 3: 112 CTAs) the same kernel runs with 6 stages instead: with 3 a CTA
// alone on its SM waited ~830 cycles per 64-channel k-atom for TMA round trips (53 us for a 6.6 us GEMM).
constexpr int PG_MAX_SEGS = 16;
constexpr int PG_MAX_CHUNKS = 2;    // 2 x 128 fp32 accumulator columns = half of TMEM

struct PGSeg { int amap, dw, dh, katoms, bmap, brow; };       // one K-segment: a tap (or the qkv "tap")
struct PGChunk { int n0, seg_begin, seg_end, kind; };          // one 128-wide accumulator

struct PGParams {
  CUtensorMap amaps[5];
  CUtensorMap bmaps[2];
  PGSeg segs[PG_MAX_SEGS];
  PGChunk chunks[PG_MAX_CHUNKS];
  int nchunks;
  int r, Wt, Ht, tiles_h, ncta;    // tile = r rows x Wt pixels of the (Ht x Wt) pixel grid this job covers; ncta = B * tiles_h
  // epilogue
  int mode;                        // 0 fprop, 1 dgrad
  void* out0; float* q; float* k; float* v;
  int out_bf16; long long y_bs;    // out0 = y (fprop: batch stride y_bs) or dx (dgrad, dense), fp32 or bf16
  int Cout, Cc, H, W, L, nh, dk, dkh, dvh, Nqkv; float qscale;           // fprop
  int Cin, Hin, Win, stride, rh, rw, pair;                                // dgrad
};

template <int PG_STAGES>
struct __align__(1024) PGSmem {
  bf16 a[PG_STAGES][128 * 64];
  bf16 b[PG_STAGES][128 * 64];
  uint64_t bar_full[PG_STAGES], bar_empty[PG_STAGES], bar_acc[PG_MAX_CHUNKS];
  uint32_t tmem_base;
};

// Several independent jobs (accumulator-chunk groups, residue classes) share one launch: blockIdx.y selects the job, so
// their CTAs fill one another's tail waves (long jobs first: blocks are dispatched in index order).
constexpr int PG_MAX_JOBS = 8;
struct PGJobs { PGParams job[PG_MAX_JOBS]; };

template <int PG_STAGES>
__global__ void __launch_bounds__(PG_THREADS, PG_STAGES <= 3 ? 2 : 1) pixel_gemm_tc_kernel(const __grid_constant__ PGJobs jobs) {
  const PGParams& p = jobs.job[blockIdx.y];
  if ((int)blockIdx.x >= p.ncta) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  PGSmem<PG_STAGES>& sm = *reinterpret_cast<PGSmem<PG_STAGES>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.tiles_h, h0 = (blockIdx.x % p.tiles_h) * p.r;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PG_STAGES; ++s) { tc::mbar_init(&sm.bar_full[s], 1); tc::mbar_init(&sm.bar_empty[s], 1); }
    for (int c = 0; c < PG_MAX_CHUNKS; ++c) tc::mbar_init(&sm.bar_acc[c], 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc<128 * PG_MAX_CHUNKS>(&sm.tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t a_bytes = (uint32_t)p.r * p.Wt * 64 * 2, b_bytes = 128 * 64 * 2;

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int c = 0; c < p.nchunks; ++c) {
        const PGChunk ch = p.chunks[c];
        for (int s = ch.seg_begin; s < ch.seg_end; ++s) {
          const PGSeg sg = p.segs[s];
          for (int a = 0; a < sg.katoms; ++a, ++it) {
            const int st = it % PG_STAGES, ph = (it / PG_STAGES) & 1;
            tc::mbar_wait(&sm.bar_empty[st], ph ^ 1);
            tc::mbar_arrive_expect_tx(&sm.bar_full[st], a_bytes + b_bytes);
            tc::tma_load_4d(sm.a[st], &p.amaps[sg.amap], &sm.bar_full[st], a * 64, sg.dw, h0 + sg.dh, b);
            tc::tma_load_2d(sm.b[st], &p.bmaps[sg.bmap], &sm.bar_full[st], a * 64, sg.brow + ch.n0);
          }
        }
      }
    }
  } else if (warp == 5) {
    // whole warp runs the uniform loop; one elected lane issues (keeps descriptors in uniform registers)
    constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 128);
    constexpr uint32_t STAGE = (128 * 128) >> 4;
    const uint32_t a_lo = tc::desc_lo_k(smem_u32(sm.a[0])), b_lo = tc::desc_lo_k(smem_u32(sm.b[0]));
    int it = 0;
    for (int c = 0; c < p.nchunks; ++c) {
      const PGChunk ch = p.chunks[c];
      uint32_t acc = 0;
      for (int s = ch.seg_begin; s < ch.seg_end; ++s) {
        const int katoms = p.segs[s].katoms;
        for (int a = 0; a < katoms; ++a, ++it) {
          const int st = it % PG_STAGES, ph = (it / PG_STAGES) & 1;
          tc::mbar_wait(&sm.bar_full[st], ph);
          tc::tc_fence_after();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::mma_ss(tmem + c * 128, tc::desc64(a_lo + st * STAGE + ks * 2), tc::desc64(b_lo + st * STAGE + ks * 2), idesc,
                         (acc | ks) ? 1u : 0u);
            tc::mma_commit(&sm.bar_empty[st]);
          }
          __syncwarp();
          acc = 1;
        }
      }
      if (tc::elect_one()) tc::mma_commit(&sm.bar_acc[c]);     // chunk c complete: its epilogue overlaps chunk c+1
      __syncwarp();
    }
  } else {
    // ===================== epilogue: thread == tile pixel == TMEM lane; chunk c as soon as its MMAs are done =====================
    // One epilogue warp per scheduler: nothing hides instruction latency, so every store is one pointer bump + STG -- the
    // job's geometry is copied to registers once (p is selected by blockIdx.y: its fields are indexed constant loads)
    // and per-channel addresses advance by a precomputed stride (profiles/r01_c_head.md: ~30 dependent instructions per
    // store made this epilogue 5x longer than the MMAs of the CTA).
    const int m = threadIdx.x;                       // 0..127
    const int Wt = p.Wt, nchunks = p.nchunks;
    const int hh = m / Wt, ww = m - hh * Wt;
    const int hrow = h0 + hh;
    const bool valid = m < p.r * Wt && hrow < p.Ht;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t rr[32];
    if (p.mode == 0) {
      const int L = p.L, Cc = p.Cc, Nqkv = p.Nqkv, dk = p.dk, dkh = p.dkh, dvh = p.dvh, nh = p.nh;
      const float qscale = p.qscale;
      const int l = hrow * p.W + ww;
      const bool vec = (dkh & 3) == 0;               // aligned groups of 4 columns never straddle a head
      const bool obf = p.out_bf16 != 0;
      float* const ybase = static_cast<float*>(p.out0) + (size_t)b * p.y_bs + l;         // fp32 y
      bf16* const ybase_h = static_cast<bf16*>(p.out0) + (size_t)b * p.y_bs + l;         // bf16 y (e.g. an autocast feature buffer)
      const size_t hs = (size_t)L * dkh;             // head stride of q / k
      float* const qbase = p.q + ((size_t)b * nh * L + l) * dkh;
      float* const kbase = p.k + ((size_t)b * nh * L + l) * dkh;
      float* const vbase = p.v + ((size_t)b * nh * L + l) * dvh;
      for (int c = 0; c < nchunks; ++c) {
        const PGChunk ch = p.chunks[c];
        tc::mbar_wait(&sm.bar_acc[c], 0);
        tc::tc_fence_after();
        for (int cb = 0; cb < 4; ++cb) {
          const int nb = ch.n0 + cb * 32;
          if (nb >= (ch.kind == 0 ? Cc : Nqkv)) break;   // uniform
          tc::tmem_ld_x32(tlane + c * 128 + cb * 32, rr);
          tc::tmem_ld_wait();
          if (!valid) continue;
          if (ch.kind == 0 && obf) {
            bf16* dst = ybase_h + (size_t)nb * L;
#pragma unroll
            for (int e = 0; e < 32; ++e, dst += L)
              if (nb + e < Cc) *dst = __float2bfloat16(__uint_as_float(rr[e]));
          } else if (ch.kind == 0) {                 // conv channels -> y NCHW (lanes = consecutive pixels: coalesced)
            float* dst = ybase + (size_t)nb * L;
            if (nb + 32 <= Cc) {
#pragma unroll
              for (int e = 0; e < 32; ++e, dst += L) *dst = __uint_as_float(rr[e]);
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e, dst += L)
                if (nb + e < Cc) *dst = __uint_as_float(rr[e]);
            }
          } else if (vec && (nb + 32 <= dk || (nb >= dk && nb + 32 <= 2 * dk))) {
            // 32 columns inside q or inside k: float4 groups walk the heads with a running (head, dim) position
            const bool isq = nb < dk;
            const int cc0 = isq ? nb : nb - dk;
            const int hd0 = cc0 / dkh;
            int ee = cc0 - hd0 * dkh;
            float* dst = (isq ? qbase : kbase) + (size_t)hd0 * hs + ee;
            const float sc = isq ? qscale : 1.f;
#pragma unroll
            for (int e4 = 0; e4 < 32; e4 += 4) {
              *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(rr[e4]) * sc, __uint_as_float(rr[e4 + 1]) * sc,
                                                            __uint_as_float(rr[e4 + 2]) * sc, __uint_as_float(rr[e4 + 3]) * sc);
              ee += 4;
              dst += 4;
              if (ee >= dkh) { ee = 0; dst += hs - dkh; }
            }
          } else {                                   // generic: q (scaled) | k | v, element by element
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int nn = nb + e;
              const float val = __uint_as_float(rr[e]);
              if (nn < dk) {
                const int hd = nn / dkh, ed = nn - hd * dkh;
                qbase[(size_t)hd * hs + ed] = val * qscale;
              } else if (nn < 2 * dk) {
                const int cc = nn - dk, hd = cc / dkh, ed = cc - hd * dkh;
                kbase[(size_t)hd * hs + ed] = val;
              } else if (nn < Nqkv) {
                const int cc = nn - 2 * dk, hd = cc / dvh, ed = cc - hd * dvh;
                vbase[(size_t)hd * L * dvh + ed] = val;
              }
            }
          }
        }
      }
    } else {
      // dgrad: chunks come in pairs (2i, 2i+1) = the two column-parity classes (rw = 0, 1) of the same 128 channels;
      // a thread owns input pixels (hi, 2*ww) and (hi, 2*ww+1) -> one 8-byte store per channel, fully coalesced rows.
      // With stride 1 (pair == 0) every chunk is its own class.
      const int Cin = p.Cin, Hin = p.Hin, Win = p.Win, stride = p.stride;
      const size_t plane = (size_t)Hin * Win;
      const int hi = hrow * stride + p.rh;
      const bool v0 = valid && hi < Hin && ww * stride + p.rw < Win;
      if (p.pair) {
        const int wi = ww * 2;
        const bool v1 = valid && hi < Hin && wi + 1 < Win;
        const bool vec2 = (Win & 1) == 0;
        const bool obf = p.out_bf16 != 0;
        float* const base = static_cast<float*>(p.out0) + (size_t)b * Cin * plane + (size_t)hi * Win + wi;
        bf16* const base_h = static_cast<bf16*>(p.out0) + (size_t)b * Cin * plane + (size_t)hi * Win + wi;
        uint32_t r2[32];
        for (int c = 0; c < nchunks; c += 2) {
          const int n0 = p.chunks[c].n0;
          tc::mbar_wait(&sm.bar_acc[c], 0);
          tc::mbar_wait(&sm.bar_acc[c + 1], 0);
          tc::tc_fence_after();
          for (int cb = 0; cb < 4; ++cb) {
            const int nb = n0 + cb * 32;
            if (nb >= Cin) break;
            tc::tmem_ld_x32(tlane + c * 128 + cb * 32, rr);
            tc::tmem_ld_x32(tlane + (c + 1) * 128 + cb * 32, r2);
            tc::tmem_ld_wait();
            if (!v0) continue;
            if (obf) {                               // bf16 dx: one 4-byte store per channel (pixel pair)
              bf16* dh = base_h + (size_t)nb * plane;
#pragma unroll
              for (int e = 0; e < 32; ++e, dh += plane) {
                if (nb + e < Cin) {
                  if (vec2) {
                    *reinterpret_cast<uint32_t*>(dh) = tc::pack_bf16x2(__uint_as_float(rr[e]), __uint_as_float(r2[e]));
                  } else {
                    dh[0] = __float2bfloat16(__uint_as_float(rr[e]));
                    if (v1) dh[1] = __float2bfloat16(__uint_as_float(r2[e]));
                  }
                }
              }
              continue;
            }
            float* dst = base + (size_t)nb * plane;
            if (vec2 && nb + 32 <= Cin) {
#pragma unroll
              for (int e = 0; e < 32; ++e, dst += plane)
                *reinterpret_cast<float2*>(dst) = make_float2(__uint_as_float(rr[e]), __uint_as_float(r2[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e, dst += plane) {
                if (nb + e < Cin) {
                  if (vec2) {
                    *reinterpret_cast<float2*>(dst) = make_float2(__uint_as_float(rr[e]), __uint_as_float(r2[e]));
                  } else {
                    dst[0] = __uint_as_float(rr[e]);
                    if (v1) dst[1] = __uint_as_float(r2[e]);
                  }
                }
              }
            }
          }
        }
      } else {
        const int wi = ww * stride + p.rw;
        const bool obf = p.out_bf16 != 0;
        float* const base = static_cast<float*>(p.out0) + (size_t)b * Cin * plane + (size_t)hi * Win + wi;
        bf16* const base_h = static_cast<bf16*>(p.out0) + (size_t)b * Cin * plane + (size_t)hi * Win + wi;
        for (int c = 0; c < nchunks; ++c) {
          const int n0 = p.chunks[c].n0;
          tc::mbar_wait(&sm.bar_acc[c], 0);
          tc::tc_fence_after();
          for (int cb = 0; cb < 4; ++cb) {
            const int nb = n0 + cb * 32;
            if (nb >= Cin) break;
            tc::tmem_ld_x32(tlane + c * 128 + cb * 32, rr);
            tc::tmem_ld_wait();
            if (!v0) continue;
            if (obf) {
              bf16* dh = base_h + (size_t)nb * plane;
#pragma unroll
              for (int e = 0; e < 32; ++e, dh += plane)
                if (nb + e < Cin) *dh = __float2bfloat16(__uint_as_float(rr[e]));
              continue;
            }
            float* dst = base + (size_t)nb * plane;
#pragma unroll
            for (int e = 0; e < 32; ++e, dst += plane)
              if (nb + e < Cin) *dst = __uint_as_float(rr[e]);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc<128 * PG_MAX_CHUNKS>(tmem);
}

// ------------------------------------------------------------------------------------------------
// packing kernels
// ------------------------------------------------------------------------------------------------
// fprop B operand: rows [t*NPc + n] = conv_w[n, :, t] (n < Cc, else 0), then rows [T*NPc + n] = qkv_w[n, :]; K = Cin.
// One thread per (n, cin): it reads the T contiguous taps of its filter element (coalesced across cin) and writes one bf16
// per tap row (coalesced across cin); no per-element index division.  grid = (ceil(CinK / 256), NPc + NPq).
__global__ void __launch_bounds__(256) pack_wf_kernel(const float* __restrict__ conv_w, const float* __restrict__ qkv_w,
                                                      bf16* __restrict__ out, int Cc, int Cin, int CinK, int T, int NPc, int Nqkv,
                                                      int NPq) {
  const int cin = blockIdx.x * blockDim.x + threadIdx.x;
  if (cin >= CinK) return;
  const int n = blockIdx.y;
  if (n < NPc) {
    const bool live = n < Cc && cin < Cin;
    const float* src = conv_w + ((size_t)n * Cin + cin) * T;
    for (int t = 0; t < T; ++t) out[((size_t)t * NPc + n) * CinK + cin] = __float2bfloat16(live ? src[t] : 0.f);
  } else {
    const int m = n - NPc;
    out[((size_t)T * NPc + m) * CinK + cin] = __float2bfloat16((m < Nqkv && cin < Cin) ? qkv_w[(size_t)m * Cin + cin] : 0.f);
  }
}

// dgrad B operands: wd rows [t*CinP + cin] = conv_w[:, cin, t] over K = co (KPc cols, zero padded);
//                   wq rows [cin] = qkv_w[:, cin] over K = n (KPq cols, zero padded)
// One thread per (cin, co): it reads the T contiguous taps of its filter element and writes one bf16 per tap row, coalesced
// along co; no per-element index division.  grid = (ceil(KPc / 128) + ceil(KPq / 128), CinP), 128 threads.
__global__ void __launch_bounds__(128) pack_wd_kernel(const float* __restrict__ conv_w, const float* __restrict__ qkv_w,
                                                      bf16* __restrict__ wd, bf16* __restrict__ wq, int Cc, int Cin, int T, int CinP,
                                                      int KPc, int Nqkv, int KPq, int nbx_conv) {
  const int cin = blockIdx.y;
  if ((int)blockIdx.x < nbx_conv) {
    const int co = blockIdx.x * 128 + threadIdx.x;
    if (co >= KPc) return;
    const bool live = co < Cc && cin < Cin;
    const float* src = conv_w + ((size_t)co * Cin + cin) * T;
    for (int t = 0; t < T; ++t) wd[((size_t)t * CinP + cin) * KPc + co] = __float2bfloat16(live ? src[t] : 0.f);
  } else {
    const int n = (blockIdx.x - nbx_conv) * 128 + threadIdx.x;
    if (n >= KPq) return;
    wq[(size_t)cin * KPq + n] = __float2bfloat16((n < Nqkv && cin < Cin) ? qkv_w[(size_t)n * Cin + cin] : 0.f);
  }
}

// dqkv (B*L, KPq) bf16 = concat(dq * qscale, dk, dv) from the head-split fp32 gradients
__global__ void pack_dqkv_kernel(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
                                 bf16* __restrict__ out, size_t pixels, int L, int nh, int dk_, int dkh, int dvh, int Nqkv,
                                 int KPq, float qscale) {
  const size_t total = pixels * KPq;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % KPq);
    const size_t pix = i / KPq;
    const int b = (int)(pix / L), l = (int)(pix - (size_t)b * L);
    float val = 0.f;
    if (n < dk_) {
      const int h = n / dkh, e = n - h * dkh;
      val = dq[((size_t)(b * nh + h) * L + l) * dkh + e] * qscale;
    } else if (n < 2 * dk_) {
      const int c = n - dk_, h = c / dkh, e = c - h * dkh;
      val = dk[((size_t)(b * nh + h) * L + l) * dkh + e];
    } else if (n < Nqkv) {
      const int c = n - 2 * dk_, h = c / dvh, e = c - h * dvh;
      val = dv[((size_t)(b * nh + h) * L + l) * dvh + e];
    }
    out[i] = __float2bfloat16(val);
  }
}

// ------------------------------------------------------------------------------------------------
// host: geometry, buffers, launches
// ------------------------------------------------------------------------------------------------
static int rows_per_tile(int Wt, int Ht) {
  int r = 128 / Wt;
  if (r > Ht) r = Ht;
  return r;
}

int tc_gemm_supported(const Dims& d) {
  if (d.stride > 2 || d.dil != 1) return fail(AACONV_E_UNSUPPORTED, "tcgen05 conv path supports stride 1|2, dilation 1");
  if (d.W > 128 || cdiv(d.Win, d.stride) > 128) return fail(AACONV_E_UNSUPPORTED, "tcgen05 conv path needs rows of <= 128 pixels");
  if (d.ks * d.ks + 1 > PG_MAX_SEGS) return fail(AACONV_E_UNSUPPORTED, "kernel_size too large for the tcgen05 conv path");
  return 0;
}

TcGemmBufs tc_gemm_bufs(const Dims& d, void* base) {
  TcGemmBufs t;
  Carver c(base);
  const int T = d.ks * d.ks;
  t.NPc = cdiv(d.Cc, 128) * 128;
  t.NPq = cdiv(d.Nqkv, 128) * 128;
  t.CinP = cdiv(d.Cin, 128) * 128;
  t.CinK = cdiv(d.Cin, 64) * 64;
  t.KPc = cdiv(d.Cout, 64) * 64;       // dy is packed with all Cout channels; weight rows past Cc are zero
  t.KPq = cdiv(d.Nqkv, 64) * 64;
  t.xh = nullptr;                        // lives in the saved block (bf16_path.cu)
  t.dyh = c.take<uint16_t>((size_t)d.B * d.L * t.KPc);
  t.dqkvh = c.take<uint16_t>((size_t)d.B * d.L * t.KPq);
  t.wf = c.take<uint16_t>(((size_t)T * t.NPc + t.NPq) * t.CinK);
  t.wd = c.take<uint16_t>((size_t)T * t.CinP * t.KPc);
  t.wq = c.take<uint16_t>((size_t)t.CinP * t.KPq);
  t.bytes = c.off;
  return t;
}

// launches the queued jobs, PG_MAX_JOBS per launch
struct PGQueue {
  std::vector<PGParams> jobs;
  void add(PGParams p, int B) { p.ncta = B * p.tiles_h; jobs.push_back(p); }
  template <int STAGES>
  int launch(const PGJobs& j, int ncta, int n, cudaStream_t st, const char* name) {
    const size_t smem = sizeof(PGSmem<STAGES>) + 1024;
    AACONV_CUDA_OK(cudaFuncSetAttribute(pixel_gemm_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pixel_gemm_tc_kernel<STAGES><<<dim3(ncta, n), PG_THREADS, smem, AACONV_ST(st)>>>(j);
    AACONV_LAUNCH_OK(name);
    return 0;
  }
  int flush(cudaStream_t st, const char* name) {
    for (size_t i = 0; i < jobs.size(); i += PG_MAX_JOBS) {
      PGJobs j;
      memset(&j, 0, sizeof j);
      const int n = (int)std::min<size_t>(PG_MAX_JOBS, jobs.size() - i);
      int ncta = 0, total = 0;
      for (int k = 0; k < n; ++k) { j.job[k] = jobs[i + k]; ncta = std::max(ncta, j.job[k].ncta); total += j.job[k].ncta; }
      if (total <= 148) AACONV_TRY(launch<6>(j, ncta, n, st, name));      // one CTA per SM: deep TMA pipeline
      else AACONV_TRY(launch<3>(j, ncta, n, st, name));                   // two CTAs per SM share the tensor pipe
    }
    jobs.clear();
    return 0;
  }
};

// channels-last 4D map over the pixels of one residue class (ph, pw) of an (B, Hs, Ws, C) tensor
static int make_nhwc_map(CUtensorMap* out, const void* base, int B, int Hs, int Ws, int C, int s, int ph, int pw, int r,
                         int Wt) {
  const int Hc = Hs > ph ? (Hs - ph + s - 1) / s : 0, Wc = Ws > pw ? (Ws - pw + s - 1) / s : 0;
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)std::max(Wc, 1), (uint64_t)std::max(Hc, 1), (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)s * C * 2, (uint64_t)s * Ws * C * 2, (uint64_t)Hs * Ws * C * 2};
  const uint32_t box[4] = {64, (uint32_t)Wt, (uint32_t)r, 1};
  const char* p = static_cast<const char*>(base) + ((size_t)ph * Ws + pw) * C * 2;
  return make_tmap_bf16(out, p, 4, dims, strides, box, nullptr);
}

static int make_2d_map(CUtensorMap* out, const void* base, int K, size_t rows) {
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)K * 2};
  const uint32_t box[2] = {64, 128};
  return make_tmap_bf16(out, base, 2, dims, strides, box, nullptr);
}

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// conv fprop + qkv projection.  y: conv channels of (B,Cout,H,W); q,k,v head-split fp32.
int tc_fprop(const Dims& d, const TcGemmBufs& t, const float* conv_w, const float* qkv_w, void* y,
             float* q, float* k, float* v, cudaStream_t st) {
  const int T = d.ks * d.ks, s = d.stride;
  pack_wf_kernel<<<dim3(cdiv(t.CinK, 256), t.NPc + t.NPq), 256, 0, AACONV_ST(st)>>>(conv_w, qkv_w, static_cast<bf16*>(t.wf), d.Cc, d.Cin,
                                                                          t.CinK, T, t.NPc, d.Nqkv, t.NPq);
  AACONV_LAUNCH_OK("pack_wf");

  PGParams p;
  memset(&p, 0, sizeof p);
  p.Wt = d.W; p.Ht = d.H; p.r = rows_per_tile(d.W, d.H); p.tiles_h = cdiv(d.H, p.r);
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw)
      AACONV_TRY(make_nhwc_map(&p.amaps[ph * s + pw], t.xh, d.B, d.Hin, d.Win, t.CinK, s, ph, pw, p.r, p.Wt));
  AACONV_TRY(make_2d_map(&p.bmaps[0], t.wf, t.CinK, (size_t)T * t.NPc + t.NPq));
  const int katoms = t.CinK / 64;
  int ns = 0;
  for (int kh = 0; kh < d.ks; ++kh)
    for (int kw = 0; kw < d.ks; ++kw) {          // input row = s*i + (kh - pad): residue + whole-pixel shift
      const int th = kh - d.pad, tw = kw - d.pad;
      const int ph = ((th % s) + s) % s, pw = ((tw % s) + s) % s;
      p.segs[ns++] = {ph * s + pw, floordiv(tw - pw, s), floordiv(th - ph, s), katoms, 0, (kh * d.ks + kw) * t.NPc};
    }
  const int seg_qkv = ns;
  p.segs[ns++] = {0, 0, 0, katoms, 0, T * t.NPc};
  p.mode = 0; p.out0 = y; p.out_bf16 = d.y_bf16; p.y_bs = d.y_bs; p.q = q; p.k = k; p.v = v;
  p.Cout = d.Cout; p.Cc = d.Cc; p.H = d.H; p.W = d.W; p.L = d.L; p.nh = d.nh; p.dk = d.dk; p.dkh = d.dkh; p.dvh = d.dvh;
  p.Nqkv = d.Nqkv; p.qscale = d.qscale;
  // accumulator chunks, up to PG_MAX_CHUNKS per launch
  // One job when everything fits the four TMEM accumulator chunks (the projection is the centre tap: its A tiles are L2-hot
  // and its scattered epilogues hide under the conv mainloop).  Otherwise groups of conv chunks (long K = 9 taps) first,
  // then groups of projection chunks (short K) whose CTAs fill the tail of the long ones -- all in one launch.
  PGQueue qu;
  const int nc_conv = cdiv(d.Cc, 128), nc_qkv = cdiv(d.Nqkv, 128);
  const int ncta = d.B * p.tiles_h;
  // accumulator chunks per CTA: as many as TMEM holds (the A tiles are then read once), but never so many that the launch
  // has fewer CTAs than SMs -- at Transition 3 (1600 pixels = 16 tiles) four chunks per CTA left 132 of 148 SMs idle
  // ... nor so many that the long conv CTAs (per x 9 taps x Cin / 64 k-atoms each) cannot be balanced over the SMs: aim for at
  // least two CTAs per SM (Transition 2 with two chunks per CTA: 64 CTAs of 144 k-atoms next to 128 of 8 -> 52 us)
  int per = PG_MAX_CHUNKS;
  while (per > 1 && ncta * (cdiv(nc_conv, per) + cdiv(nc_qkv, per)) < 2 * 148) --per;
  if (nc_conv + nc_qkv <= per) {
    p.nchunks = 0;
    for (int c = 0; c < nc_qkv; ++c) p.chunks[p.nchunks++] = PGChunk{c * 128, seg_qkv, seg_qkv + 1, 1};
    for (int c = 0; c < nc_conv; ++c) p.chunks[p.nchunks++] = PGChunk{c * 128, 0, T, 0};
    qu.add(p, d.B);
  } else {
    for (int g0 = 0; g0 < d.Cc; g0 += 128 * per) {
      p.nchunks = std::min(per, cdiv(d.Cc - g0, 128));
      for (int c = 0; c < p.nchunks; ++c) p.chunks[c] = PGChunk{g0 + c * 128, 0, T, 0};
      qu.add(p, d.B);
    }
    for (int g0 = 0; g0 < d.Nqkv; g0 += 128 * per) {
      p.nchunks = std::min(per, cdiv(d.Nqkv - g0, 128));
      for (int c = 0; c < p.nchunks; ++c) p.chunks[c] = PGChunk{g0 + c * 128, seg_qkv, seg_qkv + 1, 1};
      qu.add(p, d.B);
    }
  }
  return qu.flush(st, "conv_qkv_fprop_tc");
}

// packed backward operands shared by dgrad and wgrad: dyh (B,L,KPc) and dqkvh (B,L,KPq) = concat(dq*scale, dk, dv)
int tc_pack_grads(const Dims& d, const TcGemmBufs& t, const float* dy, const float* dq, const float* dk, const float* dv,
                  cudaStream_t st) {
  AACONV_TRY(pack_nhwc_bf16(dy, t.dyh, d.B, d.Cout, t.KPc, d.L, st));
  if (!dq) return 0;
  pack_dqkv_kernel<<<148 * 8, 256, 0, AACONV_ST(st)>>>(dq, dk, dv, static_cast<bf16*>(t.dqkvh), (size_t)d.B * d.L, d.L, d.nh, d.dk,
                                             d.dkh, d.dvh, d.Nqkv, t.KPq, d.qscale);
  AACONV_LAUNCH_OK("pack_dqkv");
  return 0;
}

// dx = conv dgrad + qkv dgrad (needs tc_pack_grads first).
int tc_dgrad(const Dims& d, const TcGemmBufs& t, const float* conv_w, const float* qkv_w, void* dx, int dx_bf16, cudaStream_t st) {
  const int T = d.ks * d.ks, s = d.stride;
  pack_wd_kernel<<<dim3(cdiv(t.KPc, 128) + cdiv(t.KPq, 128), t.CinP), 128, 0, AACONV_ST(st)>>>(
      conv_w, qkv_w, static_cast<bf16*>(t.wd), static_cast<bf16*>(t.wq), d.Cc, d.Cin, T, t.CinP, t.KPc, d.Nqkv, t.KPq, cdiv(t.KPc, 128));
  AACONV_LAUNCH_OK("pack_wd");
  // taps of the conv (and the 1x1 projection) that reach input-pixel class (rh, rw)
  auto class_segs = [&](int rh, int rw, PGSeg* out) {
    int ns = 0;
    if (d.Cc)
      for (int kh = 0; kh < d.ks; ++kh)
        for (int kw = 0; kw < d.ks; ++kw) {      // input pixel h = s*hc + rh receives dy[(h + pad - kh)/s] when divisible
          const int nh_ = rh + d.pad - kh, nw_ = rw + d.pad - kw;
          if (((nh_ % s) + s) % s || ((nw_ % s) + s) % s) continue;
          out[ns++] = {0, floordiv(nw_, s), floordiv(nh_, s), t.KPc / 64, 0, (kh * d.ks + kw) * t.CinP};
        }
    if (rh == 0 && rw == 0) out[ns++] = {1, 0, 0, t.KPq / 64, 1, 0};   // 1x1 stride-s projection touches class (0,0)
    return ns;
  };
  PGQueue qu;
  for (int rh = s - 1; rh >= 0; --rh) {           // odd rows first: they get more taps (longer jobs)
    const int Hc = d.Hin > rh ? (d.Hin - rh + s - 1) / s : 0;
    if (Hc == 0) continue;
    PGParams p;
    memset(&p, 0, sizeof p);
    PGSeg seg0[PG_MAX_SEGS], seg1[PG_MAX_SEGS];
    const int ns0 = class_segs(rh, 0, seg0), ns1 = s == 2 ? class_segs(rh, 1, seg1) : 0;
    if (s == 2 && d.Win >= 2 && ns0 > 0 && ns1 > 0 && ns0 + ns1 <= PG_MAX_SEGS) {
      // both column classes in one CTA: accumulator chunks (2i, 2i+1) = (rw 0, rw 1) of the same 128 channels
      const int Wc = (d.Win + 1) / 2;
      p.Wt = Wc; p.Ht = Hc; p.r = rows_per_tile(Wc, Hc); p.tiles_h = cdiv(Hc, p.r);
      AACONV_TRY(make_nhwc_map(&p.amaps[0], t.dyh, d.B, d.H, d.W, t.KPc, 1, 0, 0, p.r, p.Wt));
      AACONV_TRY(make_nhwc_map(&p.amaps[1], t.dqkvh, d.B, d.H, d.W, t.KPq, 1, 0, 0, p.r, p.Wt));
      AACONV_TRY(make_2d_map(&p.bmaps[0], t.wd, t.KPc, (size_t)T * t.CinP));
      AACONV_TRY(make_2d_map(&p.bmaps[1], t.wq, t.KPq, (size_t)t.CinP));
      for (int i = 0; i < ns0; ++i) p.segs[i] = seg0[i];
      for (int i = 0; i < ns1; ++i) p.segs[ns0 + i] = seg1[i];
      p.mode = 1; p.out0 = dx; p.out_bf16 = dx_bf16; p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.stride = s; p.rh = rh; p.rw = 0; p.pair = 1;
      for (int n0 = 0; n0 < d.Cin; n0 += 128 * (PG_MAX_CHUNKS / 2)) {
        const int nn = std::min(PG_MAX_CHUNKS / 2, cdiv(d.Cin - n0, 128));
        p.nchunks = 2 * nn;
        for (int c = 0; c < nn; ++c) {
          p.chunks[2 * c] = {n0 + c * 128, 0, ns0, 0};
          p.chunks[2 * c + 1] = {n0 + c * 128, ns0, ns0 + ns1, 0};
        }
        qu.add(p, d.B);
      }
      continue;
    }
    for (int rw = 0; rw < s; ++rw) {
      const int Wc = d.Win > rw ? (d.Win - rw + s - 1) / s : 0;
      if (Wc == 0) continue;
      memset(&p, 0, sizeof p);
      p.Wt = Wc; p.Ht = Hc; p.r = rows_per_tile(Wc, Hc); p.tiles_h = cdiv(Hc, p.r);
      AACONV_TRY(make_nhwc_map(&p.amaps[0], t.dyh, d.B, d.H, d.W, t.KPc, 1, 0, 0, p.r, p.Wt));
      AACONV_TRY(make_nhwc_map(&p.amaps[1], t.dqkvh, d.B, d.H, d.W, t.KPq, 1, 0, 0, p.r, p.Wt));
      AACONV_TRY(make_2d_map(&p.bmaps[0], t.wd, t.KPc, (size_t)T * t.CinP));
      AACONV_TRY(make_2d_map(&p.bmaps[1], t.wq, t.KPq, (size_t)t.CinP));
      const int ns = class_segs(rh, rw, p.segs);
      p.mode = 1; p.out0 = dx; p.out_bf16 = dx_bf16; p.Cin = d.Cin; p.Hin = d.Hin; p.Win = d.Win; p.stride = s; p.rh = rh; p.rw = rw; p.pair = 0;
      if (ns == 0) {   // no tap reaches this class (e.g. 1x1 conv with stride 2): gradient is zero there
        AACONV_TRY(tc_zero_class(d, dx, dx_bf16, rh, rw, st));
        continue;
      }
      for (int n0 = 0; n0 < d.Cin; n0 += 128 * PG_MAX_CHUNKS) {
        p.nchunks = std::min(PG_MAX_CHUNKS, cdiv(d.Cin - n0, 128));
        for (int c = 0; c < p.nchunks; ++c) p.chunks[c] = {n0 + c * 128, 0, ns, 0};
        qu.add(p, d.B);
      }
    }
  }
  return qu.flush(st, "conv_qkv_dgrad_tc");
}

__global__ void zero_class_kernel(void* __restrict__ dx, int dx_bf16, size_t planes, int Hin, int Win, int s, int rh, int rw) {
  const int Hc = (Hin - rh + s - 1) / s, Wc = (Win - rw + s - 1) / s;
  const size_t total = planes * Hc * Wc;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wc = (int)(i % Wc);
    const size_t t = i / Wc;
    const int hc = (int)(t % Hc);
    const size_t pl = t / Hc;
    store_y(dx, (pl * Hin + hc * s + rh) * Win + wc * s + rw, 0.f, dx_bf16);
  }
}

int tc_zero_class(const Dims& d, void* dx, int dx_bf16, int rh, int rw, cudaStream_t st) {
  zero_class_kernel<<<148 * 4, 256, 0, AACONV_ST(st)>>>(dx, dx_bf16, (size_t)d.B * d.Cin, d.Hin, d.Win, d.stride, rh, rw);
  AACONV_LAUNCH_OK("zero_class");
  return 0;
}

}  // namespace aaconv

// ================================================================================================
// wgrad: dW[m, n] = sum_pixels A[pixel, m] * B[pixel, n]   (A = dy or dqkv, B = x at one tap)
// Both operands are channels-last, i.e. MN-major for this contraction: a K-chunk of kr x W pixels is one TMA
// box per 64-channel atom, [pixels x 128 B], consumed through MN-major UMMA descriptors.  Split-K over pixel
// chunks; partial tiles are summed by a second kernel in a fixed order (deterministic).
// ================================================================================================
namespace aaconv {

constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 2;
constexpr int WG_KMAX = 128;        // pixels per K-chunk (rows of the smem tiles)
constexpr int WG_N = 256;           // accumulator width (input channels per task)

struct WGTap { int bmap, dw, dh; };
struct WGShape {                    // task index -> (n chunk, conv (m chunk, tap) | qkv m chunk)
  int conv_m, T, qkv_m, n_chunks, CinK;
  __host__ __device__ int per_n() const { return conv_m * T + qkv_m; }
  __host__ __device__ int ntasks() const { return per_n() * n_chunks; }
  __host__ __device__ void decode(int task, int& kind, int& m0, int& tap, int& n0, int& nn) const {
    const int nc = task / per_n(), r = task - nc * per_n();
    n0 = nc * WG_N;
    nn = min(WG_N, CinK - n0);
    if (r < conv_m * T) { kind = 0; m0 = (r / T) * 128; tap = r % T; }
    else { kind = 1; m0 = (r - conv_m * T) * 128; tap = 0; }
  }
};
struct WGParams {
  CUtensorMap amaps[2];
  CUtensorMap bmaps[4];
  WGTap taps[16];
  WGShape shape;
  int kr, Wt, kpix, chunks_per_image, nchunks_total, chunks_per_split;
  float* partial;
};

struct __align__(1024) WGSmem {
  bf16 a[WG_STAGES][2][WG_KMAX * 64];
  bf16 b[WG_STAGES][4][WG_KMAX * 64];
  uint64_t bar_full[WG_STAGES], bar_empty[WG_STAGES], bar_acc;
  uint32_t tmem_base;
};

__host__ __device__ constexpr uint32_t idesc_mn_mn(int M, int N) { return tc::idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16); }

__device__ __forceinline__ uint64_t desc_mn_tile(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WGParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WGSmem& sm = *reinterpret_cast<WGSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kind, m0, tap, n0, nn;
  p.shape.decode(blockIdx.x, kind, m0, tap, n0, nn);
  const WGTap tp = kind == 0 ? p.taps[tap] : p.taps[15];     // slot 15: the 1x1 strided projection
  const int split = blockIdx.y;
  const int g0 = split * p.chunks_per_split, g1 = min(p.nchunks_total, g0 + p.chunks_per_split);
  const int natoms_b = nn >> 6;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { tc::mbar_init(&sm.bar_full[s], 1); tc::mbar_init(&sm.bar_empty[s], 1); }
    tc::mbar_init(&sm.bar_acc, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc<256>(&sm.tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t atom_bytes = (uint32_t)p.kpix * 128;

  if (warp == 4) {
    if (lane == 0) {
      for (int g = g0, it = 0; g < g1; ++g, ++it) {
        const int st = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
        const int b = g / p.chunks_per_image, h0 = (g % p.chunks_per_image) * p.kr;
        tc::mbar_wait(&sm.bar_empty[st], ph ^ 1);
        tc::mbar_arrive_expect_tx(&sm.bar_full[st], atom_bytes * (2 + natoms_b));
        for (int a = 0; a < 2; ++a) tc::tma_load_4d(sm.a[st][a], &p.amaps[kind], &sm.bar_full[st], m0 + a * 64, 0, h0, b);
        for (int a = 0; a < natoms_b; ++a)
          tc::tma_load_4d(sm.b[st][a], &p.bmaps[tp.bmap], &sm.bar_full[st], n0 + a * 64, tp.dw, h0 + tp.dh, b);
      }
    }
  } else if (warp == 5) {
    const uint32_t idesc = idesc_mn_mn(128, nn);
    const int ksteps = p.kpix >> 4;
    constexpr uint32_t A_STAGE = (2 * WG_KMAX * 128) >> 4, B_STAGE = (4 * WG_KMAX * 128) >> 4;
    const uint32_t a_lo = tc::desc_lo_mn(smem_u32(sm.a[0][0]), WG_KMAX * 128), b_lo = tc::desc_lo_mn(smem_u32(sm.b[0][0]), WG_KMAX * 128);
    uint32_t acc = 0;
    for (int g = g0, it = 0; g < g1; ++g, ++it) {
      const int st = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
      tc::mbar_wait(&sm.bar_full[st], ph);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks)
          tc::mma_ss(tmem, tc::desc64(a_lo + st * A_STAGE + ks * 128), tc::desc64(b_lo + st * B_STAGE + ks * 128), idesc,
                     (acc | ks) ? 1u : 0u);
        tc::mma_commit(&sm.bar_empty[st]);
      }
      __syncwarp();
      acc = 1;
    }
    if (tc::elect_one()) tc::mma_commit(&sm.bar_acc);
    __syncwarp();
  } else {
    float* out = p.partial + ((size_t)(split * gridDim.x + blockIdx.x) * 128 + threadIdx.x) * WG_N;
    if (g0 < g1) {
      tc::mbar_wait(&sm.bar_acc, 0);
      tc::tc_fence_after();
      const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
      uint32_t rr[32];
      for (int cb = 0; cb < nn / 32; ++cb) {
        tc::tmem_ld_x32(tlane + cb * 32, rr);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(out + cb * 32 + e) =
              make_float4(__uint_as_float(rr[e]), __uint_as_float(rr[e + 1]), __uint_as_float(rr[e + 2]), __uint_as_float(rr[e + 3]));
      }
    } else {   // empty split: contributes zeros
      for (int e = 0; e < nn; e += 4) *reinterpret_cast<float4*>(out + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc<256>(tmem);
}

// sum the split-K partials into the parameter-gradient layouts
//   conv task (tap t, m0, n0): dWc[(co*Cin + cin)*T + t];   qkv task (m0, n0): dWq[n*Cin + cin]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, WGShape shape, int splits, float* __restrict__ dwc,
                                    float* __restrict__ dwq, int Cc, int Cin, int Nqkv) {
  const int task = blockIdx.y, ntasks = gridDim.y;
  int kind, m0, tap, n0, nn;
  shape.decode(task, kind, m0, tap, n0, nn);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 128 * WG_N; i += gridDim.x * blockDim.x) {
    const int m = m0 + i / WG_N, nl = i % WG_N, n = n0 + nl;
    if (nl >= nn || n >= Cin || m >= (kind == 0 ? Cc : Nqkv)) continue;
    // the partials are summed in split order (deterministic); loads are issued four at a time so that the chain is
    // bound by the adds, not by one memory round trip per split
    const float* src = partial + (size_t)task * 128 * WG_N + i;
    const size_t step = (size_t)ntasks * 128 * WG_N;
    float s = 0.f;
    int sp = 0;
    for (; sp + 4 <= splits; sp += 4) {
      const float a = src[(size_t)sp * step], b = src[(size_t)(sp + 1) * step], c = src[(size_t)(sp + 2) * step],
                  d = src[(size_t)(sp + 3) * step];
      s = (((s + a) + b) + c) + d;
    }
    for (; sp < splits; ++sp) s += src[(size_t)sp * step];
    if (kind == 0) dwc[((size_t)m * Cin + n) * shape.T + tap] = s;
    else dwq[(size_t)m * Cin + n] = s;
  }
}

static int wgrad_kr(int W) {
  int best = 0;
  for (int kr = 1; kr * W <= WG_KMAX; ++kr)
    if ((kr * W) % 16 == 0) best = kr;
  return best;
}

int tc_wgrad_supported(const Dims& d) {
  if (tc_gemm_supported(d)) return AACONV_E_UNSUPPORTED;
  if (wgrad_kr(d.W) == 0) return fail(AACONV_E_UNSUPPORTED, "no pixel chunk of W=%d rows is a multiple of 16 within 128", d.W);
  return 0;
}

// split-K over pixel chunks only while the tasks alone leave SMs idle.  (Measured at Transition 3, 156 tasks: splitting 4-8 ways
// to even out the 1.05 waves made wgrad 62 -> 87 us and its reduction 27 -> 54 us: the kernel is bound by its partial-tile
// traffic, not by the tail wave.)
static int wgrad_splits(int nt, int nchunks_total) { return std::max(1, std::min(148 / nt, nchunks_total)); }

size_t tc_wgrad_partial_floats(const Dims& d) {
  const int T = d.ks * d.ks, nch = cdiv(d.Cin, WG_N);
  const int ntasks = (cdiv(d.Cc, 128) * T + cdiv(d.Nqkv, 128)) * nch;
  const int kr = wgrad_kr(d.W);
  const int nchunks_total = kr ? d.B * cdiv(d.H, kr) : 1;
  return (size_t)ntasks * wgrad_splits(ntasks, nchunks_total) * 128 * WG_N;
}

// xh, dyh (and dqkvh when dwq != NULL) must already hold the packed operands of this step.
int tc_wgrad(const Dims& d, const TcGemmBufs& t, float* dwc, float* dwq, float* partial, cudaStream_t st) {
  const int T = d.ks * d.ks, s = d.stride;
  WGParams p;
  memset(&p, 0, sizeof p);
  p.kr = wgrad_kr(d.W); p.Wt = d.W; p.kpix = p.kr * d.W;
  p.chunks_per_image = cdiv(d.H, p.kr); p.nchunks_total = d.B * p.chunks_per_image;
  AACONV_TRY(make_nhwc_map(&p.amaps[0], t.dyh, d.B, d.H, d.W, t.KPc, 1, 0, 0, p.kr, p.Wt));
  AACONV_TRY(make_nhwc_map(&p.amaps[1], t.dqkvh, d.B, d.H, d.W, t.KPq, 1, 0, 0, p.kr, p.Wt));
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw)
      AACONV_TRY(make_nhwc_map(&p.bmaps[ph * s + pw], t.xh, d.B, d.Hin, d.Win, t.CinK, s, ph, pw, p.kr, p.Wt));
  for (int kh = 0; kh < d.ks; ++kh)
    for (int kw = 0; kw < d.ks; ++kw) {
      const int th = kh - d.pad, tw = kw - d.pad;
      const int ph = ((th % s) + s) % s, pw = ((tw % s) + s) % s;
      p.taps[kh * d.ks + kw] = {ph * s + pw, floordiv(tw - pw, s), floordiv(th - ph, s)};
    }
  p.taps[15] = {0, 0, 0};
  p.shape.conv_m = (dwc && d.Cc) ? cdiv(d.Cc, 128) : 0;
  p.shape.T = T;
  p.shape.qkv_m = dwq ? cdiv(d.Nqkv, 128) : 0;
  p.shape.n_chunks = cdiv(d.Cin, WG_N);
  p.shape.CinK = t.CinK;
  const int nt = p.shape.ntasks();
  if (nt == 0) return 0;
  int splits = wgrad_splits(nt, p.nchunks_total);
  while (splits > 1 && (size_t)nt * splits * 128 * WG_N > tc_wgrad_partial_floats(d)) --splits;   // the caller sized `partial` for all tasks
  p.chunks_per_split = cdiv(p.nchunks_total, splits);
  splits = cdiv(p.nchunks_total, p.chunks_per_split);
  p.partial = partial;
  const size_t smem = sizeof(WGSmem) + 1024;
  AACONV_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_tc_kernel<<<dim3(nt, splits), WG_THREADS, smem, AACONV_ST(st)>>>(p);
  AACONV_LAUNCH_OK("conv_qkv_wgrad_tc");
  wgrad_reduce_kernel<<<dim3(std::max(16, std::min(128, 2368 / nt)), nt), 256, 0, AACONV_ST(st)>>>(partial, p.shape, splits, dwc, dwq, d.Cc, d.Cin, d.Nqkv);
  AACONV_LAUNCH_OK("wgrad_reduce");
  return 0;
}

}  // namespace aaconv
