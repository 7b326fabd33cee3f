// fp16 hi/lo fused attention, "two teams" form: the production kernel of OG_PREC_FP16X3.
//
// Same arithmetic, operands and TMEM layout as csrc/attention_f16.cuh; what changes is who does the softmax.  With fp16
// operands a key block costs the tensor pipe 768 cycles (24 MMAs), but ONE group of softmax warps needs ~1650 cycles per
// block for its serial chain (S load -> row max exchange -> 32 exponentials + hi/lo split -> P store -> fold of O_{i-1};
// event trace profiles/r02_trace_attention_f16_v1.txt) - the tensor pipe idled half of the time.  Here TWO teams of two
// softmax warpgroups work on ALTERNATE key blocks: team e owns every block with (global index & 1) == e together with the
// S/P/O TMEM buffers e, keeps its own online-softmax state (row max, row sum, 32 output channels per thread) and the two
// partial results of a tile are merged once, at the end of the tile, through shared memory:
//     out = (acc_0 2^(m_0 - m) + acc_1 2^(m_1 - m)) / (l_0 2^(m_0 - m) + l_1 2^(m_1 - m)),   m = max(m_0, m_1)
// Within a team the two warpgroups split a block by columns exactly as before (half-row maxima through shared memory and
// one 256-thread named barrier).  The element-wise chain uses the packed fp32x2 instructions (FFMA2 / FADD2): one issue
// slot for two logits.  20 warps: 0-7 team 0, 8-15 team 1, 16 TMA, 17 QK^T issue + TMEM, 18 P.V issue; setmaxnreg moves the
// producer warps' registers to the softmax warpgroups (104 registers: no spills; with the accumulators spilled the kernel ran at half
// the speed - 256 KB of local-memory traffic per key block through the same 128 B/clk port as shared memory).
// Measured (self layer, 32 x 4 x 2048 x 2048): 0.360 ms = 381 TF/s against 0.417 ms for the one-team form.
#pragma once
#include "tc_common.cuh"
#include "attention_f16.cuh"     // F16AttnScales, TcAttnArgs
#include <math_constants.h>
#include <stdlib.h>
#include <algorithm>

// OG_ATTN_EARLY_S (default 0: parity-clean on B200 but measured SLOWER, 0.336 -> 0.344 ms per self layer): the QK^T issuer waits for "S_{i-2} is in the team's registers" (s_free, arrived right after the
// tcgen05.ld of the logits) instead of "P_{i-2} has been written" (p_full, ~1300 cycles later).  The event trace of the p_full form
// (profiles/r02_trace_attention_f16_two_teams.txt) shows QK^T_{i+2} issued the moment P_i is handed over and its logits seen ~830
// cycles after that: a team's cycle was softmax (1650) + MMA round trip (830) per two key blocks.  S and P occupy disjoint TMEM
// columns, so the next QK^T of a buffer only has to wait for the load.  The wait for S disappears - and the kernel slows down,
// presumably because the teams are no longer held in anti-phase by the issue order (profiles/README.md, finding 8).
#ifndef OG_ATTN_EARLY_S
#define OG_ATTN_EARLY_S 0
#endif
// OG_ATTN_FOLD_EARLY (default 0: measured neutral): fold O_{i-2} into the registers BEFORE the exponentials of block i (in four
// 8-column loads: the logits are live) instead of after P_i has been handed over, so that P.V_i never waits for the fold.
#ifndef OG_ATTN_FOLD_EARLY
#define OG_ATTN_FOLD_EARLY 0
#endif
// OG_ATTN_MERGER_LAST (default 1; kernel template parameter SWAP, chosen per launch: see attention_f16t_launch_t): the team that owns a tile's LAST key block merges and stores the tile; the other
// team deposits its partial result (bar.arrive, no wait) and starts the next tile's first block, which is the earlier one.
// With the merge fixed on team 0 and an even block count, team 0 waited half a cycle for team 1's last block, merged (~2500
// cycles) and only then turned to a block whose logits had been ready all along (~6000 cycles per tile boundary in the trace).  Measured -1.5 %.
#ifndef OG_ATTN_MERGER_LAST
#define OG_ATTN_MERGER_LAST 1
#endif
// OG_ATTN_PAIR_BAR (default 1: -0.4 %): the half-row maxima are exchanged between the two warps that own the same 32 rows
// (one in each warpgroup of the team) behind a 64-thread named barrier of their own instead of the team's 256-thread barrier:
// a pair no longer waits for the slowest of the team's eight warps on every key block.
#ifndef OG_ATTN_PAIR_BAR
#define OG_ATTN_PAIR_BAR 1
#endif

namespace og {
namespace tcat {
constexpr int BM = 128, BNK = 64, DH = 64, HD = 32;
constexpr int MAX_STAGES = 4;
template <int CG> __host__ __device__ constexpr int stages() { return CG == 2 ? 4 : 2; }      // CG = 1 stages whole K / V tiles: 32 KB per stage
constexpr int THREADS = 640;
constexpr int TMEM_COLS = 512;
constexpr int COL_QHI = 0, COL_QLO = 32, COL_SP = 64, COL_O = 320;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float P_SHIFT = 14.f;
constexpr int REGS_SOFTMAX = 104, REGS_PRODUCER = 64;        // setmaxnreg moves registers INSIDE the CTA's launch allocation:
                                                             // 512 x 104 + 128 x 64 = 61440 = 640 threads x 96 registers (checked by the launcher)

struct __align__(16) Barriers {
  uint64_t k_full[MAX_STAGES], k_empty[MAX_STAGES], v_full[MAX_STAGES], v_empty[MAX_STAGES];
  uint64_t q_ready, q_free, s_full[2], p_full[2], o_full[2], o_empty[2], s_free[2];
  uint32_t tmem_base;
};
template <int CG> __host__ __device__ constexpr int k_stage_bytes() { return 2 * (BNK / CG) * 128; }
template <int CG> __host__ __device__ constexpr int v_stage_bytes() { return 2 * (DH / CG) * 128; }
constexpr int XCH_FLOATS = 2 * 2 * 2 * 128;                  // [team][block parity][warpgroup][row] half-row maxima
constexpr int LM_FLOATS = 2 * 2 * 2 * 128 * 2;               // [tile parity][team][warpgroup][row] (mc, l)
constexpr int MRG_STRIDE = HD + 1;
constexpr int MRG_FLOATS = 2 * 2 * 128 * MRG_STRIDE;         // [tile parity][warpgroup][row][33] team 1's partial output
constexpr int QST_FLOATS = 16 * 32 * HD;                     // per softmax warp a [32 rows x 32 channels] staging tile (Q in, O out)
template <int CG> __host__ __device__ constexpr int smem_bytes() {
  return 1024 + stages<CG>() * (k_stage_bytes<CG>() + v_stage_bytes<CG>()) + 512 + (XCH_FLOATS + LM_FLOATS + MRG_FLOATS + QST_FLOATS) * 4;
}

__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long fsub2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
}  // namespace tcat

template <int CG, int SWAP>
__global__ void __launch_bounds__(tcat::THREADS, 1) attention_f16t_kernel(const __grid_constant__ CUtensorMap map_khi,
                                                                          const __grid_constant__ CUtensorMap map_klo,
                                                                          const __grid_constant__ CUtensorMap map_vhi,
                                                                          const __grid_constant__ CUtensorMap map_vlo,
                                                                          TcAttnArgs a, F16AttnScales sc) {
  using namespace tcat;
  using namespace tc;
  constexpr int KROWS = BNK / CG, VCH = DH / CG;
  constexpr int K_HALF = KROWS * 128, V_HALF = VCH * 128;
  constexpr int STAGES = stages<CG>();

  launch_dependents();
  extern __shared__ uint8_t og_tcat_smem_raw[];
  uint8_t* smem = tc::align_smem_1024(og_tcat_smem_raw);
  uint8_t* sK = smem;
  uint8_t* sV = smem + STAGES * k_stage_bytes<CG>();
  Barriers* bars = reinterpret_cast<Barriers*>(sV + STAGES * v_stage_bytes<CG>());
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  float* lm = xch + XCH_FLOATS;
  float* mrg = lm + LM_FLOATS;
  float* qst_all = mrg + MRG_FLOATS;

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
    mbar_init(&bars->q_ready, 8 * CG);               // the eight warps of the team that hands a tile's Q over (in both CTAs)
    mbar_init(&bars->q_free, 1);
    for (int j = 0; j < 2; ++j) {                    // buffer j belongs to team j: eight warps per CTA
      mbar_init(&bars->s_full[j], 1); mbar_init(&bars->p_full[j], 8 * CG);
      mbar_init(&bars->o_full[j], 1); mbar_init(&bars->o_empty[j], 8 * CG);
      mbar_init(&bars->s_free[j], 8 * CG);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_khi); prefetch_tensormap(&map_klo);
    prefetch_tensormap(&map_vhi); prefetch_tensormap(&map_vlo);
  }
  if (CG == 2) cluster_sync_all();
  if (warp == 17) { if (CG == 2) tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base); else tmem_alloc<TMEM_COLS>(&bars->tmem_base); }
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

  if (warp >= 16) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
  if (warp == 16) {
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
        mbar_wait_t(&bars->k_empty[s], ph ^ 1);
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
        mbar_wait_t(&bars->v_empty[s], ph ^ 1);
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
  } else if ((warp == 17 || warp == 18) && crank == 0) {
    // ------------------------------------------------------------------ MMA issuers (leader CTA only when paired)
    const uint32_t idesc_qk = make_idesc_f16(BM * CG, BNK);
    const uint32_t idesc_pv = make_idesc_f16(BM * CG, DH);
    auto mma = [&](uint32_t d, uint32_t at, uint64_t bd, uint32_t id, uint32_t acc) {
      if (CG == 2) umma_f16_ts_pair(d, at, bd, id, acc); else umma_f16_ts(d, at, bd, id, acc);
    };
    if (warp == 17) {
      int it = 0, nt = 0;
      for (int t = t_first; t < a.ntiles; t += t_stride, ++nt) {
      mbar_wait_t(&bars->q_ready, nt & 1);
      for (int iloc = 0; iloc < nblk; ++iloc, ++it) {
        const int i = it;
        const int s = i % STAGES, ph = (i / STAGES) & 1, j = i & 1;
        OG_TRACE_EVT(0, i);
        mbar_wait_t(&bars->k_full[s], ph);
#if OG_ATTN_EARLY_S
        if (i >= 2) mbar_wait_t(&bars->s_free[j], ((i - 2) >> 1) & 1);     // team j holds S_{i-2} in registers (S and P columns are disjoint)
#else
        if (i >= 2) mbar_wait_t(&bars->p_full[j], ((i - 2) >> 1) & 1);     // team j has read S_{i-2} out of its buffer
#endif
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
        mbar_wait_t(&bars->v_full[s], ph);
        mbar_wait_t(&bars->p_full[j], jph);
        OG_TRACE_EVT(3, i);
        mbar_wait_t(&bars->o_empty[j], jph ^ 1);
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
    // ------------------------------------------------------------------ softmax teams
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
    const int team = warp >> 3;                      // owns key blocks with (global index & 1) == team and the TMEM buffers `team`
    const int g = (warp >> 2) & 1;                   // column half inside the team: logit columns / output channels [32g, 32g+32)
    const int qd = warp & 3;
    const int trow = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float* xch_t = xch + team * (2 * 2 * 128);
    const float s_q = f16_scale_for(__ldcg(sc.q_amax));
    const float s_k = __ldcg(sc.k_scale), s_v = __ldcg(sc.v_scale);
    const float c1 = a.scale * LOG2E / (s_q * s_k);
    const unsigned long long c1_2 = pack2(c1, c1);
    const float inv_sv = 1.f / s_v;
    float omax = 0.f;
    constexpr int LPR = HD / 4, RPI = 32 / LPR;      // 8 lanes fetch one row's 128 bytes
    float* qst = qst_all + warp * (32 * HD);         // this warp's staging tile [32][32] (16-byte chunks XOR-swizzled by row)
    const uint32_t sp = tmem + lane_base + COL_SP + 128 * team;
    const uint32_t o_addr = tmem + lane_base + COL_O + 64 * team + g * HD;
    uint64_t* const bar_s = &bars->s_full[team];
    uint64_t* const bar_p = &bars->p_full[team];
    uint64_t* const bar_of = &bars->o_full[team];
    uint64_t* const bar_oe = &bars->o_empty[team];
#if !OG_ATTN_PAIR_BAR
    const int bar_id = 1 + team;
#endif

    auto load_q = [&](const TilePos& p) {            // cp.async: lands during the tile
      const int r_in = lane / LPR, ch = lane % LPR;
#pragma unroll
      for (int k = 0; k < LPR; ++k) {
        const int row = k * RPI + r_in, grow = tile_q0(p) + qd * 32 + row;
        float* dst = qst + row * HD + ((ch ^ (row & (LPR - 1))) * 4);
        const float* src = a.q + (int64_t)p.b * a.strideq + (int64_t)grow * a.ldq + p.h * DH + g * HD + ch * 4;
        if (grow < a.nq) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto write_q = [&]() {                           // staged Q row -> scale, split, pack -> TMEM
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(qst + lane * HD + ((c ^ (lane & (LPR - 1))) * 4));
        split_f16x2(v.x * s_q, v.y * s_q, hi[2 * c], lo[2 * c]);
        split_f16x2(v.z * s_q, v.w * s_q, hi[2 * c + 1], lo[2 * c + 1]);
      }
      __syncwarp();
      tmem_st_32x16(tmem + lane_base + COL_QHI + g * 16, hi);
      tmem_st_32x16(tmem + lane_base + COL_QLO + g * 16, lo);
      tmem_wait_st();
      tc_fence_before();
      arrive_leader(&bars->q_ready);
    };
    TilePos tp = tile_pos(t_first);
    if (t_first < a.ntiles && team == 0) load_q(tp);
    int it0 = 0, nt = 0;                             // global index of the tile's first key block, tiles done

#pragma unroll 1
    for (int t = t_first; t < a.ntiles; t += t_stride, ++nt, it0 += nblk) {
    const int h = tp.h, b = tp.b;
    const int grow = tile_q0(tp) + trow;
    tp = tile_next(tp);
    const bool has_next = t + t_stride < a.ntiles;
    // the team that owns the tile's LAST key block sees the last QK^T retire first: it hands the next tile's Q over
    const bool writer_next = has_next && (((it0 + nblk - 1) & 1) == team);
    if (nt == 0 && team == 0) write_q();
    if (writer_next) load_q(tp);

    unsigned long long acc2[HD / 2];
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) acc2[c] = 0ull;
    float m_run = -CUDART_INF_F, mc_run = -CUDART_INF_F, l_run = 0.f, corr_prev = 0.f;
    int prev = -1;

    auto fold_o = [&](int i, float corr) {           // acc = acc * corr + my channels of O_i
      mbar_wait_t(bar_of, (i >> 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32(o_addr, o);
      tmem_wait_ld();
      const unsigned long long corr2 = pack2(corr, corr);
#pragma unroll
      for (int c = 0; c < 16; ++c) acc2[c] = ffma2(acc2[c], corr2, pack2(__uint_as_float(o[2 * c]), __uint_as_float(o[2 * c + 1])));
      tc_fence_before();
      arrive_leader(bar_oe);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(7, i);
    };

#pragma unroll 1
    for (int iloc = (team - it0) & 1; iloc < nblk; iloc += 2) {
      const int i = it0 + iloc;
      const int par = (i >> 1) & 1;
      const int kbase = iloc * BNK + 32 * g;
      mbar_wait_t(bar_s, par);
      tc_fence_after();
      if (warp == 0 && lane == 0) OG_TRACE_EVT(5, i);
      if (iloc == nblk - 1 && writer_next) {         // the tile's last QK^T has retired: the next tile's Q goes in now
        mbar_wait_t(&bars->q_free, nt & 1);
        tc_fence_after();
        write_q();
      }
      uint32_t s[32];
      tmem_ld_32x32(sp + 32 * g, s);
      tmem_wait_ld();
#if OG_ATTN_EARLY_S
      tc_fence_before();
      arrive_leader(&bars->s_free[team]);            // QK^T_{i+2} may overwrite the S columns from here on
#endif
      if (kbase + 32 > a.nk) {
#pragma unroll
        for (int c = 0; c < 32; ++c) if (kbase + c >= a.nk) s[c] = __float_as_uint(-CUDART_INF_F);
      }
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 32; c += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 1])));
      xch_t[(par * 2 + g) * 128 + trow] = mx;
#if OG_ATTN_PAIR_BAR
      asm volatile("bar.sync %0, 64;" ::"r"(4 + team * 4 + qd) : "memory");       // only the two warps that share these 32 rows
#else
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
#endif
      mx = fmaxf(mx, xch_t[(par * 2 + (g ^ 1)) * 128 + trow]);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(8, i);
      const float m_new = fmaxf(m_run, mx);
      const float mc = fmaf(m_new, c1, -P_SHIFT);
      const float corr = ex2_approx(mc_run - mc);
      const unsigned long long nmc2 = pack2(-mc, -mc);
      unsigned long long rs2 = 0ull;
#if OG_ATTN_FOLD_EARLY
      if (prev >= 0) {                                 // P_i goes where P_prev sits: the team's previous P.V must have read it;
        mbar_wait_t(bar_of, (prev >> 1) & 1);            // its result is folded right here, 8 columns at a time
        tc_fence_after();
        const unsigned long long corr2 = pack2(corr_prev, corr_prev);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t o[8];
          tmem_ld_32x8(o_addr + 8 * q, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 4; ++c)
            acc2[4 * q + c] = ffma2(acc2[4 * q + c], corr2, pack2(__uint_as_float(o[2 * c]), __uint_as_float(o[2 * c + 1])));
        }
        tc_fence_before();
        arrive_leader(bar_oe);
      }
#else
      if (prev >= 0) {                                 // P_i goes where P_prev sits: the team's previous P.V must have read it
        mbar_wait_t(bar_of, (prev >> 1) & 1);            // (issued a block and a half ago: no wait in steady state; folded below)
        tc_fence_after();
      }
#endif
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {                 // two halves of 16 columns: P leaves the registers as soon as it is split
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          float t0, t1;
          unpack2(ffma2(pack2(__uint_as_float(s[16 * hh + c]), __uint_as_float(s[16 * hh + c + 1])), c1_2, nmc2), t0, t1);
          const float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
          const unsigned long long p2 = pack2(p0, p1);
          rs2 = fadd2(rs2, p2);
          const __half2 hh2 = __floats2half2_rn(p0, p1);
          const float2 hf = __half22float2(hh2);
          float l0, l1;
          unpack2(fsub2(p2, pack2(hf.x, hf.y)), l0, l1);
          const __half2 ll = __floats2half2_rn(l0, l1);
          hi[c >> 1] = *reinterpret_cast<const uint32_t*>(&hh2);
          lo[c >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        tmem_st_32x8(sp + 64 + 16 * g + 8 * hh, hi);   // P_hi: keys [32g + 16hh, +16) = packed columns [16g + 8hh, +8)
        tmem_st_32x8(sp + 96 + 16 * g + 8 * hh, lo);   // (this thread has waited for the team's previous P.V in fold_o)
      }
      if (warp == 0 && lane == 0) OG_TRACE_EVT(9, i);
      tmem_wait_st();
      tc_fence_before();
      arrive_leader(bar_p);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(6, i);
      float r0, r1;
      unpack2(rs2, r0, r1);
      l_run = fmaf(l_run, corr, r0 + r1);
      m_run = m_new; mc_run = mc;
#if !OG_ATTN_FOLD_EARLY
      if (prev >= 0) fold_o(prev, corr_prev);
#endif
      prev = i; corr_prev = corr;
    }
    if (prev >= 0) fold_o(prev, corr_prev);

    // ---- merge the two teams' partial results of this tile (depositor -> shared memory -> merger), normalise, store
    const int tp2 = nt & 1;
    const int merger = SWAP ? ((it0 + nblk - 1) & 1) : 0;     // SWAP: the team that finishes the tile LAST merges; the other one deposits and moves on
    float* lm_t = lm + tp2 * (2 * 2 * 128 * 2);
    reinterpret_cast<float2*>(lm_t)[(team * 2 + g) * 128 + trow] = make_float2(mc_run, l_run);
    if (team != merger) {
      float* mr = mrg + ((tp2 * 2 + g) * 128 + trow) * MRG_STRIDE;
#pragma unroll
      for (int c = 0; c < 16; ++c) { float x, y; unpack2(acc2[c], x, y); mr[2 * c] = x; mr[2 * c + 1] = y; }
    }
    if (SWAP && team != merger) {                    // producer side of the named barrier: no wait (PTX bar.arrive / bar.sync pattern)
      __threadfence_block();
      asm volatile("bar.arrive 3, 512;" ::: "memory");
    } else {
      asm volatile("bar.sync 3, 512;" ::: "memory");
    }
    if (team == merger) {
      const float2* lmv = reinterpret_cast<const float2*>(lm_t);
      const float2 a0 = lmv[(team * 2 + (g ^ 1)) * 128 + trow], b0 = lmv[((team ^ 1) * 2 + 0) * 128 + trow], b1 = lmv[((team ^ 1) * 2 + 1) * 128 + trow];
      const float mc0 = mc_run, mc1 = b0.x;            // the two warpgroups of a team share their running maximum
      const float mcf = fmaxf(mc0, mc1);
      const float f0 = ex2_approx(mc0 - mcf), f1 = ex2_approx(mc1 - mcf);     // a team without blocks: 2^(-inf) = 0
      const float l_tot = fmaf(l_run + a0.y, f0, (b0.y + b1.y) * f1);
      const float inv = inv_sv / l_tot;
      const float w0 = f0 * inv, w1 = f1 * inv;
      const float* mr = mrg + ((tp2 * 2 + g) * 128 + trow) * MRG_STRIDE;
      float out[HD];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float x, y; unpack2(acc2[c], x, y);
        out[2 * c] = fmaf(x, w0, mr[2 * c] * w1); out[2 * c + 1] = fmaf(y, w0, mr[2 * c + 1] * w1);
      }
      if (sc.out_amax && grow < a.nq) {
#pragma unroll
        for (int c = 0; c < HD; ++c) omax = fmaxf(omax, fabsf(out[c]));
      }
      // row-coalesced stores through this warp's staging tile (free: the next tile's Q, if this team loads it, is issued below)
#pragma unroll
      for (int c = 0; c < HD / 4; ++c)
        *reinterpret_cast<float4*>(qst + lane * HD + ((c ^ (lane & (LPR - 1))) * 4)) = make_float4(out[4 * c], out[4 * c + 1], out[4 * c + 2], out[4 * c + 3]);
      __syncwarp();
      const int r_in = lane / LPR, ch = lane % LPR;
#pragma unroll
      for (int k = 0; k < LPR; ++k) {
        const int row = k * RPI + r_in, orow_g = grow - lane + row;
        if (orow_g < a.nq)
          *reinterpret_cast<float4*>(a.out + (int64_t)b * a.strideo + (int64_t)orow_g * a.ldo + h * DH + g * HD + ch * 4) =
              *reinterpret_cast<const float4*>(qst + row * HD + ((ch ^ (row & (LPR - 1))) * 4));
      }
      __syncwarp();
    }
    }
    if (sc.out_amax) {                               // both teams may have merged tiles
      omax = warp_max(omax);
      if (lane == 0 && omax > 0.f) atomic_amax(sc.out_amax, omax);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 17) { tc_fence_after(); if (CG == 2) tmem_dealloc_pair<tcat::TMEM_COLS>(tmem); else tmem_dealloc<tcat::TMEM_COLS>(tmem); }
}

template <int CG, int SWAP>
inline int attention_f16t_launch_ts(const TcAttnArgs& a, const F16AttnScales& sc, const __half* khi, const __half* klo, int64_t ldk,
                                   const __half* vthi, const __half* vtlo, int64_t ldvt, cudaStream_t stream) {
  using namespace tcat;
  CUtensorMap mkh, mkl, mvh, mvl;
  int rc;
  if ((rc = tc::make_tmap_2d_f16(&mkh, khi, (uint64_t)a.batch * a.nk, (uint64_t)a.d, (uint64_t)ldk, BNK / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d_f16(&mkl, klo, (uint64_t)a.batch * a.nk, (uint64_t)a.d, (uint64_t)ldk, BNK / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d_f16(&mvh, vthi, (uint64_t)a.batch * a.d, (uint64_t)a.nk, (uint64_t)ldvt, DH / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d_f16(&mvl, vtlo, (uint64_t)a.batch * a.d, (uint64_t)a.nk, (uint64_t)ldvt, DH / CG)) != OG_OK) return rc;
  static DeviceFlags attr_set;
  if (attr_set.once()) {
    OG_CUDA(cudaFuncSetAttribute(attention_f16t_kernel<CG, SWAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<CG>()));
  }
  {  // the register hand-over only works if the compiled register count gives the CTA the pool the two setmaxnreg values add up to
    static int regs_ok = -1;
    if (regs_ok < 0) {
      cudaFuncAttributes fa;
      OG_CUDA(cudaFuncGetAttributes(&fa, attention_f16t_kernel<CG, SWAP>));
      regs_ok = (fa.numRegs * THREADS >= 512 * REGS_SOFTMAX + 128 * REGS_PRODUCER) ? 1 : 0;
    }
    if (!regs_ok) return fail(OG_EUNSUPPORTED, "attention_f16t: register pool too small for the setmaxnreg split (rebuild)");
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
  OG_CUDA(cudaLaunchKernelEx(&cfg, attention_f16t_kernel<CG, SWAP>, mkh, mkl, mvh, mvl, ap, sc));
  launch_counter()++;
  return OG_OK;
}

// The role swap at the end of a tile (OG_ATTN_MERGER_LAST) pairs a waiting `bar.sync` with a non-waiting `bar.arrive` on one named
// barrier, which is only sound while no team can reach the barrier for tile t+1 before the other team has reached it for tile t (one
// generation would then be completed by two arrivals of the same team).  QK^T is issued in global block order and QK^T(g) waits for
// P(g-2): a team's third block of tile t+1 therefore sits behind a QK^T that waits for the OTHER team's first P of tile t+1, which
// that team writes after tile t's barrier - from five key blocks per tile on the order is a hard dependency, below that it would rest
// on timing alone (tests/test_attention_protocol_model.py explores every interleaving of this protocol).  Sequences of fewer than six
// key blocks therefore run the symmetric form: team 0 merges, both teams wait.
template <int CG>
inline int attention_f16t_launch_t(const TcAttnArgs& a, const F16AttnScales& sc, const __half* khi, const __half* klo, int64_t ldk,
                                   const __half* vthi, const __half* vtlo, int64_t ldvt, cudaStream_t stream) {
  const int nblk = (a.nk + tcat::BNK - 1) / tcat::BNK;
  if (OG_ATTN_MERGER_LAST && nblk >= 6) return attention_f16t_launch_ts<CG, 1>(a, sc, khi, klo, ldk, vthi, vtlo, ldvt, stream);
  return attention_f16t_launch_ts<CG, 0>(a, sc, khi, klo, ldk, vthi, vtlo, ldvt, stream);
}

}  // namespace og
