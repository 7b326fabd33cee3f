// tcgen05 fused multi-head softmax attention with fp16 hi/lo operands ("3xFP16", kind::f16): the fp32-grade contract of
// csrc/attention_tc.cuh (three products per contraction, long reductions folded in registers with RN adds) at TWICE the
// MMA rate and half the K / V bytes.  Replaces softmax_attention (reference models/superglue/attention.py:8-19).
//
//   S_i  = Q . K_i^T      A = Q hi/lo packed halves resident in TMEM (scaled by sQ, split once per tile),
//                         B = K_i hi/lo fp16 tiles [keys x 64 channels] (128-byte swizzled rows, TMA), scale sK
//   P_i  = 2^14 exp(S_i scale - m_i)   one query row per thread, written back to TMEM as packed halves hi/lo
//                                      (2^14: the row maximum of P is 1; its hi/lo pair then resolves 2^-39)
//   O_i  = P_i . V_i      B = V_i^T hi/lo fp16 tiles [64 channels x 64 keys], scale sV; fresh fp32 accumulator per block
//   acc  = acc * exp(m_{i-1} - m_i) + O_i     in registers of the softmax threads; out = acc / (l sV)
// sQ derives from the Q tensor's tracked amax (q_amax), sK / sV are the scales the projection GEMM wrote K / V^T with
// (csrc/linear_f16.cuh); they fold into the softmax's single FFMA constant and the final normalisation.
//
// Structure as attention_tc.cuh: persistent CTAs (pairs: cta_group::2, M = 256), 12 warps: warpgroups 0 / 1 = softmax /
// correction (column split: WG g owns logit columns [32g, 32g+32) and output channels [32g, 32g+32)), warp 8 = TMA, warps
// 9 / 10 = MMA issuers for QK^T / P.V.  TMEM: Q_hi [0,32) Q_lo [32,64) | S/P buffer j at 64 + 128 j: S [0,64) P_hi [64,96)
// P_lo [96,128) | O_j at 320 + 64 j.  P no longer shares columns with S, so QK^T_{i+2} only waits for the softmax to have
// READ S_i (p_full), not for P.V_i.
#pragma once
#include "tc_common.cuh"
#include "attention_tc.cuh"      // TcAttnArgs, attention_tc_pair_mode
#include <math_constants.h>
#include <stdlib.h>
#include <algorithm>

namespace og {

struct F16AttnScales {
  const float* q_amax;           // device: max |Q| (tracked by the projection GEMM)
  const float* k_scale;          // device: scale K hi/lo were written with
  const float* v_scale;          // device: scale V^T hi/lo were written with
  float* out_amax;               // optional: max |out| (atomicMax; zeroed by the caller)
  int swap_halves;               // debug probe of the packed TMEM operand layout
};

namespace tcaf {
constexpr int BM = 128, BNK = 64, DH = 64;
constexpr int STAGES = 4;
constexpr int THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr int COL_QHI = 0, COL_QLO = 32, COL_SP = 64, COL_O = 320;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float P_SHIFT = 14.f;                 // P is written as 2^14 p

struct __align__(16) Barriers {
  uint64_t k_full[STAGES], k_empty[STAGES], v_full[STAGES], v_empty[STAGES];
  uint64_t q_ready, q_free, s_full[2], p_full[2], o_full[2], o_empty[2];
  uint32_t tmem_base;
};
template <int CG> __host__ __device__ constexpr int k_stage_bytes() { return 2 * (BNK / CG) * 128; }        // hi + lo, 128-byte rows
template <int CG> __host__ __device__ constexpr int v_stage_bytes() { return 2 * (DH / CG) * 128; }
template <int CG> __host__ __device__ constexpr int stage_bytes() { return CG == 2 ? 2 * 8 * 32 * (DH / 2) * 4 : 0; }
template <int CG> __host__ __device__ constexpr int smem_bytes() {
  return 1024 + STAGES * (k_stage_bytes<CG>() + v_stage_bytes<CG>()) + 512 + 8 * 128 * 4 + stage_bytes<CG>();
}
}  // namespace tcaf

template <int CG>
__global__ void __launch_bounds__(tcaf::THREADS, 1) attention_f16_kernel(const __grid_constant__ CUtensorMap map_khi,
                                                                         const __grid_constant__ CUtensorMap map_klo,
                                                                         const __grid_constant__ CUtensorMap map_vhi,
                                                                         const __grid_constant__ CUtensorMap map_vlo,
                                                                         TcAttnArgs a, F16AttnScales sc) {
  using namespace tcaf;
  using namespace tc;
  constexpr int KROWS = BNK / CG, VCH = DH / CG;
  constexpr int K_HALF = KROWS * 128;             // bytes of the hi (or lo) part of a K stage
  constexpr int V_HALF = VCH * 128;

  launch_dependents();
  extern __shared__ uint8_t og_tcaf_smem_raw[];
  uint8_t* smem = tc::align_smem_1024(og_tcaf_smem_raw);
  uint8_t* sK = smem;
  uint8_t* sV = smem + STAGES * k_stage_bytes<CG>();
  Barriers* bars = reinterpret_cast<Barriers*>(sV + STAGES * v_stage_bytes<CG>());

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;
  const int nblk = (a.nk + BNK - 1) / BNK;
  const int t_first = blockIdx.x / CG, t_stride = gridDim.x / CG;
  struct TilePos { int qg, h, b; };
  auto tile_pos = [&](int t) { TilePos p; p.qg = t % a.nqg; p.h = (t / a.nqg) % a.num_heads; p.b = t / (a.nqg * a.num_heads); return p; };
  const TilePos t_step = tile_pos(t_stride);
  auto tile_next = [&](TilePos p) {
    p.qg += t_step.qg; if (p.qg >= a.nqg) { p.qg -= a.nqg; ++p.h; }
    p.h += t_step.h;   if (p.h >= a.num_heads) { p.h -= a.num_heads; ++p.b; }
    p.b += t_step.b;
    return p;
  };
  auto tile_q0 = [&](const TilePos& p) { return (p.qg * CG + (int)crank) * BM; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bars->k_full[i], 1); mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1); mbar_init(&bars->v_empty[i], 1);
    }
    mbar_init(&bars->q_ready, 8 * CG);
    mbar_init(&bars->q_free, 1);
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bars->s_full[j], 1); mbar_init(&bars->p_full[j], 8 * CG);
      mbar_init(&bars->o_full[j], 1); mbar_init(&bars->o_empty[j], 8 * CG);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_khi); prefetch_tensormap(&map_klo);
    prefetch_tensormap(&map_vhi); prefetch_tensormap(&map_vlo);
  }
  if (CG == 2) cluster_sync_all();
  if (warp == 9) { if (CG == 2) tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base); else tmem_alloc<TMEM_COLS>(&bars->tmem_base); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  grid_dependency_wait();
  auto arrive_leader = [&](uint64_t* bar) {
    __syncwarp();
    if (lane == 0) { if (CG == 1 || crank == 0) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); }
  };
  auto commit = [&](uint64_t* bar) { if (CG == 2) umma_commit_pair(bar); else umma_commit(bar); };

  if (warp >= 8) {
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      TilePos tp = tile_pos(t_first);
      for (int t = t_first; t < a.ntiles; t += t_stride, tp = tile_next(tp)) {
      const int h = tp.h, b = tp.b;
      const int krow0 = b * a.nk;
      const int vrow = b * a.d + h * DH;
      for (int i = 0; i < nblk; ++i, ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&bars->k_empty[s], ph ^ 1);
        if (crank == 0) mbar_arrive_expect_tx(&bars->k_full[s], CG * k_stage_bytes<CG>());
        uint8_t* kd = sK + s * k_stage_bytes<CG>();
        const int kr = krow0 + i * BNK + (int)crank * KROWS;
        if (CG == 2) {
          tma_load_2d_pair(kd, &map_khi, &bars->k_full[s], h * DH, kr);
          tma_load_2d_pair(kd + K_HALF, &map_klo, &bars->k_full[s], h * DH, kr);
        } else {
          tma_load_2d(kd, &map_khi, &bars->k_full[s], h * DH, kr);
          tma_load_2d(kd + K_HALF, &map_klo, &bars->k_full[s], h * DH, kr);
        }
        mbar_wait(&bars->v_empty[s], ph ^ 1);
        if (crank == 0) mbar_arrive_expect_tx(&bars->v_full[s], CG * v_stage_bytes<CG>());
        uint8_t* vd = sV + s * v_stage_bytes<CG>();
        const int vr = vrow + (int)crank * VCH;
        if (CG == 2) {
          tma_load_2d_pair(vd, &map_vhi, &bars->v_full[s], i * BNK, vr);
          tma_load_2d_pair(vd + V_HALF, &map_vlo, &bars->v_full[s], i * BNK, vr);
        } else {
          tma_load_2d(vd, &map_vhi, &bars->v_full[s], i * BNK, vr);
          tma_load_2d(vd + V_HALF, &map_vlo, &bars->v_full[s], i * BNK, vr);
        }
      }
      }
    }
  } else if ((warp == 9 || warp == 10) && crank == 0) {
    // ------------------------------------------------------------------ MMA issuers (leader CTA only when paired)
    const uint32_t idesc_qk = make_idesc_f16(BM * CG, BNK);
    const uint32_t idesc_pv = make_idesc_f16(BM * CG, DH);
    auto mma = [&](uint32_t d, uint32_t at, uint64_t bd, uint32_t id, uint32_t acc) {
      if (CG == 2) umma_f16_ts_pair(d, at, bd, id, acc); else umma_f16_ts(d, at, bd, id, acc);
    };
    if (warp == 9) {
      int it = 0, nt = 0;
      for (int t = t_first; t < a.ntiles; t += t_stride, ++nt) {
      mbar_wait(&bars->q_ready, nt & 1);
      for (int iloc = 0; iloc < nblk; ++iloc, ++it) {
        const int i = it;
        const int s = i % STAGES, ph = (i / STAGES) & 1, j = i & 1;
        OG_TRACE_EVT(0, i);
        mbar_wait(&bars->k_full[s], ph);
        if (i >= 2) mbar_wait(&bars->p_full[j], ((i - 2) >> 1) & 1);     // the softmax has read S_{i-2} out of this buffer
        tc_fence_after();
        OG_TRACE_EVT(1, i);
        if (elect_one()) {
          const uint32_t khi = smem_u32(sK + s * k_stage_bytes<CG>()), klo = khi + K_HALF;
          const uint32_t d_s = tmem + COL_SP + 128 * j;
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            const uint64_t dhi = make_sdesc_sw128(khi + kk * 32), dlo = make_sdesc_sw128(klo + kk * 32);
            mma(d_s, tmem + COL_QLO + kk * 8, dhi, idesc_qk, kk ? 1u : 0u);
            mma(d_s, tmem + COL_QHI + kk * 8, dlo, idesc_qk, 1u);
            mma(d_s, tmem + COL_QHI + kk * 8, dhi, idesc_qk, 1u);
          }
          commit(&bars->k_empty[s]);
          commit(&bars->s_full[j]);
          if (iloc == nblk - 1) commit(&bars->q_free);
        }
        __syncwarp();
      }
      }
    } else {
      const int ntot = nblk * ((a.ntiles - t_first + t_stride - 1) / t_stride);
      for (int i = 0; i < ntot; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1, j = i & 1, jph = (i >> 1) & 1;
        OG_TRACE_EVT(2, i);
        mbar_wait(&bars->v_full[s], ph);
        mbar_wait(&bars->p_full[j], jph);
        OG_TRACE_EVT(3, i);
        mbar_wait(&bars->o_empty[j], jph ^ 1);
        tc_fence_after();
        OG_TRACE_EVT(4, i);
        if (elect_one()) {
          const uint32_t vhi = smem_u32(sV + s * v_stage_bytes<CG>()), vlo = vhi + V_HALF;
          const uint32_t p_hi = tmem + COL_SP + 128 * j + 64, p_lo = p_hi + 32;
          const uint32_t d_o = tmem + COL_O + 64 * j;
#pragma unroll
          for (int kk = 0; kk < BNK / 16; ++kk) {
            const uint64_t dhi = make_sdesc_sw128(vhi + kk * 32), dlo = make_sdesc_sw128(vlo + kk * 32);
            mma(d_o, p_lo + kk * 8, dhi, idesc_pv, kk ? 1u : 0u);
            mma(d_o, p_hi + kk * 8, dlo, idesc_pv, 1u);
            mma(d_o, p_hi + kk * 8, dhi, idesc_pv, 1u);
          }
          commit(&bars->v_empty[s]);
          commit(&bars->o_full[j]);
        }
        __syncwarp();
      }
    }
  }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    constexpr int HD = DH / 2;                       // output channels per warpgroup (32)
    const int g = warp >> 2;
    const int qd = warp & 3;
    const int trow = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float* xch = reinterpret_cast<float*>(bars + 1);
    // operand scales (device scalars written by the producing kernels; read after the grid dependency wait)
    const float s_q = f16_scale_for(__ldcg(sc.q_amax));
    const float s_k = __ldcg(sc.k_scale), s_v = __ldcg(sc.v_scale);
    const float c1 = a.scale * LOG2E / (s_q * s_k);  // logits arrive multiplied by sQ sK
    const float inv_sv = 1.f / s_v;
    int it = 0, nt = 0;
    float4 qv[HD / 4];
    float omax = 0.f;
    constexpr bool STAGED = CG == 2;
    constexpr int LPR = HD / 4, RPI = 32 / LPR;
    float* qst = xch + 8 * 128 + warp * (2 * 32 * HD);
    float* ost = qst + 32 * HD;
    auto load_q = [&](const TilePos& p) {
      const int h = p.h, b = p.b;
      if constexpr (STAGED) {
        const int r_in = lane / LPR, ch = lane % LPR;
#pragma unroll
        for (int k = 0; k < LPR; ++k) {
          const int row = k * RPI + r_in, grow = tile_q0(p) + qd * 32 + row;
          float* dst = qst + row * HD + ((ch ^ (row & (LPR - 1))) * 4);
          const float* src = a.q + (int64_t)b * a.strideq + (int64_t)grow * a.ldq + h * DH + g * HD + ch * 4;
          if (grow < a.nq) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
          else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      } else {
        const int grow = tile_q0(p) + trow;
        const float* qrow = a.q + (int64_t)b * a.strideq + (int64_t)grow * a.ldq + h * DH + g * HD;
#pragma unroll
        for (int c = 0; c < HD / 4; ++c)
          qv[c] = (grow < a.nq) ? __ldg(reinterpret_cast<const float4*>(qrow + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto write_q = [&]() {                           // qv -> scale, split, pack -> TMEM (A operand of every QK^T of a tile)
      if constexpr (STAGED) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) qv[c] = *reinterpret_cast<const float4*>(qst + lane * HD + ((c ^ (lane & (LPR - 1))) * 4));
        __syncwarp();
      }
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        split_f16x2(qv[c].x * s_q, qv[c].y * s_q, hi[2 * c], lo[2 * c]);
        split_f16x2(qv[c].z * s_q, qv[c].w * s_q, hi[2 * c + 1], lo[2 * c + 1]);
      }
      if (sc.swap_halves) {
#pragma unroll
        for (int w = 0; w < 16; ++w) { hi[w] = __byte_perm(hi[w], 0, 0x1032); lo[w] = __byte_perm(lo[w], 0, 0x1032); }
      }
      tmem_st_32x16(tmem + lane_base + COL_QHI + g * 16, hi);
      tmem_st_32x16(tmem + lane_base + COL_QLO + g * 16, lo);
      tmem_wait_st();
      tc_fence_before();
      arrive_leader(&bars->q_ready);
    };
    TilePos tp = tile_pos(t_first);
    if (t_first < a.ntiles) load_q(tp);

#pragma unroll 1
    for (int t = t_first; t < a.ntiles; t += t_stride, ++nt) {
    const int h = tp.h, b = tp.b;
    const int grow = tile_q0(tp) + trow;
    const bool row_ok = grow < a.nq;
    tp = tile_next(tp);
    if (nt == 0) write_q();
    if (t + t_stride < a.ntiles) load_q(tp);

    float acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    float m_run = -CUDART_INF_F, mc_run = -CUDART_INF_F, l_run = 0.f, corr_prev = 0.f;

    auto fold_o = [&](int i, float corr) {
      const int j = i & 1, jph = (i >> 1) & 1;
      mbar_wait(&bars->o_full[j], jph);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32(tmem + lane_base + COL_O + 64 * j + g * HD, o);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = fmaf(acc[c], corr, __uint_as_float(o[c]));
      tc_fence_before();
      arrive_leader(&bars->o_empty[j]);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(7, i);
    };

#pragma unroll 1
    for (int iloc = 0; iloc < nblk; ++iloc, ++it) {
      const int i = it;
      const int j = i & 1, jph = (i >> 1) & 1;
      const uint32_t sp = tmem + lane_base + COL_SP + 128 * j;
      const int kbase = iloc * BNK + 32 * g;
      mbar_wait(&bars->s_full[j], jph);
      tc_fence_after();
      if (warp == 0 && lane == 0) OG_TRACE_EVT(5, i);
      if (iloc == nblk - 1 && t + t_stride < a.ntiles) {
        mbar_wait(&bars->q_free, nt & 1);
        tc_fence_after();
        write_q();
      }
      uint32_t s[32];
      tmem_ld_32x32(sp + 32 * g, s);
      tmem_wait_ld();
      if (kbase + 32 > a.nk) {
#pragma unroll
        for (int c = 0; c < 32; ++c) if (kbase + c >= a.nk) s[c] = __float_as_uint(-CUDART_INF_F);
      }
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 32; c += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 1])));
      xch[(j * 2 + g) * 128 + trow] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // also: BOTH warpgroups have read S_i (nothing below touches S)
      mx = fmaxf(mx, xch[(j * 2 + (g ^ 1)) * 128 + trow]);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(8, i);
      const float m_new = fmaxf(m_run, mx);
      const float mc = fmaf(m_new, c1, -P_SHIFT);      // p = 2^(s c1 - mc) = 2^14 exp(scale (s - m))
      const float corr = ex2_approx(mc_run - mc);
      float r0 = 0.f, r1 = 0.f;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), c1, -mc));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), c1, -mc));
        r0 += p0; r1 += p1;
        split_f16x2(p0, p1, hi[c >> 1], lo[c >> 1]);
      }
      if (sc.swap_halves) {
#pragma unroll
        for (int w = 0; w < 16; ++w) { hi[w] = __byte_perm(hi[w], 0, 0x1032); lo[w] = __byte_perm(lo[w], 0, 0x1032); }
      }
      if (warp == 0 && lane == 0) OG_TRACE_EVT(9, i);
      tmem_st_32x16(sp + 64 + 16 * g, hi);             // P_hi: keys [32g, 32g+32) = packed columns [16g, 16g+16)
      tmem_st_32x16(sp + 96 + 16 * g, lo);             // (this thread has waited for P.V_{i-2} in fold_o(i-2))
      tmem_wait_st();
      tc_fence_before();
      arrive_leader(&bars->p_full[j]);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(6, i);
      l_run = fmaf(l_run, corr, r0 + r1);
      m_run = m_new; mc_run = mc;
      if (iloc >= 1) fold_o(i - 1, corr_prev);
      corr_prev = corr;
    }
    fold_o(it - 1, corr_prev);

    float* xl = xch + 4 * 128 + (nt & 1) * 256;
    xl[g * 128 + trow] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = inv_sv / (l_run + xl[(g ^ 1) * 128 + trow]);
    if (sc.out_amax && row_ok) {
#pragma unroll
      for (int c = 0; c < HD; ++c) omax = fmaxf(omax, fabsf(acc[c] * inv));
    }
    if constexpr (STAGED) {
#pragma unroll
      for (int c = 0; c < HD / 4; ++c)
        *reinterpret_cast<float4*>(ost + lane * HD + ((c ^ (lane & (LPR - 1))) * 4)) =
            make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
      __syncwarp();
      const int r_in = lane / LPR, ch = lane % LPR;
#pragma unroll
      for (int k = 0; k < LPR; ++k) {
        const int row = k * RPI + r_in, orow_g = grow - lane + row;
        if (orow_g < a.nq)
          *reinterpret_cast<float4*>(a.out + (int64_t)b * a.strideo + (int64_t)orow_g * a.ldo + h * DH + g * HD + ch * 4) =
              *reinterpret_cast<const float4*>(ost + row * HD + ((ch ^ (row & (LPR - 1))) * 4));
      }
      __syncwarp();
    } else if (row_ok) {
      float* orow = a.out + (int64_t)b * a.strideo + (int64_t)grow * a.ldo + h * DH + g * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 4)
        *reinterpret_cast<float4*>(orow + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
    }
    }
    if (sc.out_amax) {
      omax = warp_max(omax);
      if (lane == 0 && omax > 0.f) atomic_amax(sc.out_amax, omax);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 9) { tc_fence_after(); if (CG == 2) tmem_dealloc_pair<tcaf::TMEM_COLS>(tmem); else tmem_dealloc<tcaf::TMEM_COLS>(tmem); }
}

// khi/klo: fp16 [batch*nk, ldk];  vthi/vtlo: fp16 [batch*d, ldvt]  (ld in elements, multiples of 8)
template <int CG>
inline int attention_f16_launch_t(const TcAttnArgs& a, const F16AttnScales& sc, const __half* khi, const __half* klo, int64_t ldk,
                                  const __half* vthi, const __half* vtlo, int64_t ldvt, cudaStream_t stream) {
  using namespace tcaf;
  CUtensorMap mkh, mkl, mvh, mvl;
  int rc;
  if ((rc = tc::make_tmap_2d_f16(&mkh, khi, (uint64_t)a.batch * a.nk, (uint64_t)a.d, (uint64_t)ldk, BNK / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d_f16(&mkl, klo, (uint64_t)a.batch * a.nk, (uint64_t)a.d, (uint64_t)ldk, BNK / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d_f16(&mvh, vthi, (uint64_t)a.batch * a.d, (uint64_t)a.nk, (uint64_t)ldvt, DH / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d_f16(&mvl, vtlo, (uint64_t)a.batch * a.d, (uint64_t)a.nk, (uint64_t)ldvt, DH / CG)) != OG_OK) return rc;
  static DeviceFlags attr_set;
  if (attr_set.once()) {
    OG_CUDA(cudaFuncSetAttribute(attention_f16_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<CG>()));
  }
  TcAttnArgs ap = a;
  ap.nqg = cdiv(cdiv(a.nq, BM), CG);
  ap.ntiles = ap.nqg * a.num_heads * a.batch;
  const int sms = device_info().ok ? device_info().sm_count : 148;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(std::min(ap.ntiles, sms / CG) * CG);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem_bytes<CG>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = tc::pdl_mode() ? 2 : 1;
  OG_CUDA(cudaLaunchKernelEx(&cfg, attention_f16_kernel<CG>, mkh, mkl, mvh, mvl, ap, sc));
  launch_counter()++;
  return OG_OK;
}

inline bool attention_f16_eligible(int head_dim, int64_t ldq, int64_t ldk, int64_t ldvt, int64_t ldo) {
  return head_dim == 64 && ldq % 4 == 0 && ldk % 8 == 0 && ldvt % 8 == 0 && ldo % 4 == 0;
}

inline int attention_f16_launch(const TcAttnArgs& a, const F16AttnScales& sc, const __half* khi, const __half* klo, int64_t ldk,
                                const __half* vthi, const __half* vtlo, int64_t ldvt, int head_dim, cudaStream_t stream) {
  if (head_dim != 64) return fail(OG_EUNSUPPORTED, "attention_f16: head_dim %d != 64", head_dim);
  if (!sc.q_amax || !sc.k_scale || !sc.v_scale) return fail(OG_EINVAL, "attention_f16: operand scales missing");
  return attention_tc_pair_mode() != 0 ? attention_f16_launch_t<2>(a, sc, khi, klo, ldk, vthi, vtlo, ldvt, stream)
                                       : attention_f16_launch_t<1>(a, sc, khi, klo, ldk, vthi, vtlo, ldvt, stream);
}

}  // namespace og
