// tcgen05 fused multi-head softmax attention with fp32-grade accuracy ("3xTF32").
// Replaces softmax_attention (reference models/superglue/attention.py:8-19); the N x M probability
// tensor never exists outside TMEM.
//
// Persistent: one CTA per SM loops over tiles; a tile = 128 queries of one (batch, head); key blocks of 64.
// (TMEM allocation, barrier set-up, pipeline fill and the output epilogue of one tile overlap the next tile's
// loads and MMAs; as one-tile CTAs these cost ~17% of the kernel.)
//   S_i  = Q . K_i^T      A = Q hi/lo resident in TMEM (split once), B = K_i hi/lo tiles (TMA, smem)
//   P_i  = exp(S_i*scale - m_i)   by 128 softmax threads, one query row each (row max / sum are
//                                 thread-local: no shuffles), written back to TMEM split hi/lo
//   O_i  = P_i . V_i      A = P_i hi/lo in TMEM, B = V_i^T hi/lo tiles (TMA, smem), fresh accumulator
//   acc  = acc * exp(m_{i-1} - m_i) + O_i      in registers of the softmax threads
// Every contraction is three tf32 MMAs (lo.hi + hi.lo + hi.hi) with fp32 accumulation in TMEM.
// Warp roles (12 warps): warpgroups 0 and 1 = softmax / correction, splitting every key block by columns (WG g:
// logit columns [32g, 32g+32) and output channels [g*Dh/2, (g+1)*Dh/2); the half-row maxima are exchanged through
// smem once per block), warp 8 = TMA producer, warps 9 / 10 = MMA issuers for QK^T / P.V (warp 9 also owns TMEM).
//
// Operand layouts in HBM (written by the projection GEMM's epilogue, csrc/linear_tc.cuh):
//   Q        fp32  [rows, ldq]            keypoint-major, head h = columns [h*Dh, (h+1)*Dh)
//   K hi/lo  tf32  [batch*nk, ldk]        keypoint-major
//   Vt hi/lo tf32  [batch*d, ldvt]        channel-major (the reference's own [B, d, M] layout)
#pragma once
#include "tc_common.cuh"
#include <math_constants.h>
#include <stdlib.h>
#include <algorithm>

namespace og {

struct TcAttnArgs {
  const float* q; int64_t ldq, strideq;        // strideq: floats between batch items
  float* out; int64_t ldo, strideo;
  int batch, nq, nk, num_heads, d;
  float scale;
  int nqg, ntiles;                              // query-block groups per sequence, total tiles (set by the launcher)
};

namespace tca {
constexpr int BM = 128, BNK = 64;               // queries per CTA, keys per block
constexpr int STAGES = 3;
constexpr int THREADS = 384;
constexpr int TMEM_COLS = 512;
// TMEM columns:  Q_hi [0,64)  Q_lo [64,128)  SP_j [128+128j, +128) = {S / P_hi: 64, P_lo: 64}  O_j [384+64j, +64)
constexpr int COL_QHI = 0, COL_QLO = 64, COL_SP = 128, COL_O = 384;
constexpr float LOG2E = 1.4426950408889634f;

struct __align__(16) Barriers {      // 16: the staging tiles behind the exchange buffer are accessed as float4
  uint64_t k_full[STAGES], k_empty[STAGES], v_full[STAGES], v_empty[STAGES];
  uint64_t q_ready, q_free, s_full[2], p_full[2], o_full[2], o_empty[2], plo_free[2];
  uint32_t tmem_base;
};
// CG = 1: one CTA per 128 queries.  CG = 2: a CTA pair (cta_group::2) per 256 queries; every MMA spans both SMs
// (M = 256) and each CTA stages only half of the K rows / V^T channels of a block.
template <int DH, int CG> __host__ __device__ constexpr int k_stage_bytes() { return 2 * (BNK / CG) * DH * 4; }        // hi + lo
template <int DH, int CG> __host__ __device__ constexpr int v_stage_bytes() { return 2 * (DH / CG) * BNK * 4; }
// CG = 2 has shared memory to spare: per softmax warp one [32 rows x Dh/2] staging tile for the next tile's Q rows and one for the
// output rows, so that global loads / stores move whole 128-byte rows per instruction (a thread owns a row - TMEM lane - and
// its own 128 bytes: 32 different cache lines per warp instruction, which cost ~2000 LSU cycles per tile each way)
template <int DH, int CG> __host__ __device__ constexpr int stage_bytes() { return CG == 2 ? 2 * 8 * 32 * (DH / 2) * 4 : 0; }
template <int DH, int CG> __host__ __device__ constexpr int smem_bytes() {
  return 1024 + (STAGES * (k_stage_bytes<DH, CG>() + v_stage_bytes<DH, CG>()) < 36864 ? 36864 : STAGES * (k_stage_bytes<DH, CG>() + v_stage_bytes<DH, CG>())) + 512 + 8 * 128 * 4 +
         stage_bytes<DH, CG>();
}
}  // namespace tca

template <int DH, int CG>
__global__ void __launch_bounds__(tca::THREADS, 1) attention_tc_kernel(const __grid_constant__ CUtensorMap map_khi,
                                                                       const __grid_constant__ CUtensorMap map_klo,
                                                                       const __grid_constant__ CUtensorMap map_vhi,
                                                                       const __grid_constant__ CUtensorMap map_vlo,
                                                                       TcAttnArgs a) {
  using namespace tca;
  using namespace tc;
  static_assert(DH == 32 || DH == 64, "head_dim 32 or 64");
  constexpr int KROWS = BNK / CG, VCH = DH / CG;  // K rows / V^T channels of a block staged by THIS CTA
  constexpr int KBLK = KROWS * 128;               // bytes of one [KROWS keys x 32 ch] K block
  constexpr int VBLK = VCH * 128;                 // bytes of one [VCH ch x 32 keys] V^T block
  constexpr int K_HALF = (DH / 32) * KBLK;        // hi (or lo) part of a K stage
  constexpr int V_HALF = 2 * VBLK;

  launch_dependents();                                       // the next kernel may take this SM as soon as this CTA leaves it
  extern __shared__ uint8_t og_tca_smem_raw[];
  uint8_t* smem = tc::align_smem_1024(og_tca_smem_raw);
  uint8_t* sK = smem;
  uint8_t* sV = smem + STAGES * k_stage_bytes<DH, CG>();
  Barriers* bars = reinterpret_cast<Barriers*>(sV + STAGES * v_stage_bytes<DH, CG>());

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform (setmaxnreg)
  const int lane = threadIdx.x & 31;
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;          // 0 = leader of the pair
  const int nblk = (a.nk + BNK - 1) / BNK;
  const int t_first = blockIdx.x / CG, t_stride = gridDim.x / CG;      // tile list of this CTA (pair)
  // tile t -> (query-block group, head, batch); query groups fastest so that co-running CTAs share K / V in L2.
  // The decomposition is advanced incrementally from tile to tile (t += t_stride): the integer divisions run once per
  // kernel instead of three times per tile - once-per-tile code is instruction-cache-cold (ncu: ~130 instruction-cache
  // misses per tile and SM; the tile hand-over cost ~9 % of a 32-block tile), so it is kept short.
  struct TilePos { int qg, h, b; };
  auto tile_pos = [&](int t) { TilePos p; p.qg = t % a.nqg; p.h = (t / a.nqg) % a.num_heads; p.b = t / (a.nqg * a.num_heads); return p; };
  const TilePos t_step = tile_pos(t_stride);          // (b may exceed the batch here: it is only ever added)
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
    mbar_init(&bars->q_ready, 8 * CG);               // one arrival per softmax warp (8 per CTA); CG = 2: the leader also counts the peer's
    mbar_init(&bars->q_free, 1);                     // every QK^T of the current tile has retired: Q may be replaced
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bars->s_full[j], 1); mbar_init(&bars->p_full[j], 8 * CG);
      mbar_init(&bars->o_full[j], 1); mbar_init(&bars->o_empty[j], 8 * CG); mbar_init(&bars->plo_free[j], 1);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_khi); prefetch_tensormap(&map_klo);
    prefetch_tensormap(&map_vhi); prefetch_tensormap(&map_vlo);
  }
  if (CG == 2) cluster_sync_all();                 // both CTAs' barriers exist before anything signals them
  if (warp == 9) { if (CG == 2) tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base); else tmem_alloc<TMEM_COLS>(&bars->tmem_base); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  grid_dependency_wait();                                    // nothing above touches memory written by the previous kernel
  // signal a barrier that lives in the leader CTA (local arrive for CG = 1 / the leader itself)
  // (one lane per warp, after every lane of the warp has completed and fenced its own TMEM accesses)
  auto arrive_leader = [&](uint64_t* bar) {
    __syncwarp();
    if (lane == 0) { if (CG == 1 || crank == 0) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); }
  };
  auto commit = [&](uint64_t* bar) { if (CG == 2) umma_commit_pair(bar); else umma_commit(bar); };

  if (warp >= 8) {
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;                                   // key blocks loaded so far (ring position), across tiles
      TilePos tp = tile_pos(t_first);
      for (int t = t_first; t < a.ntiles; t += t_stride, tp = tile_next(tp)) {
      const int h = tp.h, b = tp.b;
      const int krow0 = b * a.nk;                   // K rows of this batch item
      const int vrow = b * a.d + h * DH;            // V^T rows (channels) of this (batch, head)
      for (int i = 0; i < nblk; ++i, ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&bars->k_empty[s], ph ^ 1);
        if (crank == 0) mbar_arrive_expect_tx(&bars->k_full[s], CG * k_stage_bytes<DH, CG>());   // both CTAs' bytes land on the leader's barrier
        uint8_t* kd = sK + s * k_stage_bytes<DH, CG>();
        const int kr = krow0 + i * BNK + (int)crank * KROWS;          // my half of the block's keys
#pragma unroll
        for (int cb = 0; cb < DH / 32; ++cb) {
          if (CG == 2) {
            tma_load_2d_pair(kd + cb * KBLK, &map_khi, &bars->k_full[s], h * DH + cb * 32, kr);
            tma_load_2d_pair(kd + K_HALF + cb * KBLK, &map_klo, &bars->k_full[s], h * DH + cb * 32, kr);
          } else {
            tma_load_2d(kd + cb * KBLK, &map_khi, &bars->k_full[s], h * DH + cb * 32, kr);
            tma_load_2d(kd + K_HALF + cb * KBLK, &map_klo, &bars->k_full[s], h * DH + cb * 32, kr);
          }
        }
        mbar_wait(&bars->v_empty[s], ph ^ 1);
        if (crank == 0) mbar_arrive_expect_tx(&bars->v_full[s], CG * v_stage_bytes<DH, CG>());
        uint8_t* vd = sV + s * v_stage_bytes<DH, CG>();
        const int vr = vrow + (int)crank * VCH;                       // my half of the head's channels
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          if (CG == 2) {
            tma_load_2d_pair(vd + kb * VBLK, &map_vhi, &bars->v_full[s], i * BNK + kb * 32, vr);
            tma_load_2d_pair(vd + V_HALF + kb * VBLK, &map_vlo, &bars->v_full[s], i * BNK + kb * 32, vr);
          } else {
            tma_load_2d(vd + kb * VBLK, &map_vhi, &bars->v_full[s], i * BNK + kb * 32, vr);
            tma_load_2d(vd + V_HALF + kb * VBLK, &map_vlo, &bars->v_full[s], i * BNK + kb * 32, vr);
          }
        }
      }
      }
    }
  } else if ((warp == 9 || warp == 10) && crank == 0) {
    // ------------------------------------------------------------------ MMA issuers (leader CTA only when paired)
    // Two issuing warps: measured (scripts/trace_attn.py), ONE thread needs ~40-47 cycles per tcgen05.mma while an
    // M128 N64 K8 tf32 MMA occupies the tensor pipe for 32 - the issuing thread, not the pipe, paced the kernel.
    // Warp 9 issues every QK^T, warp 10 every P.V; they only meet through mbarriers.
    const uint32_t idesc_qk = make_idesc_tf32(BM * CG, BNK);
    const uint32_t idesc_pv = make_idesc_tf32(BM * CG, DH);
    auto mma = [&](uint32_t d, uint32_t at, uint64_t bd, uint32_t id, uint32_t acc) {
      if (CG == 2) umma_tf32_ts_pair(d, at, bd, id, acc); else umma_tf32_ts(d, at, bd, id, acc);
    };
    if (warp == 9) {
      int it = 0, nt = 0;                             // key blocks / tiles done so far
      for (int t = t_first; t < a.ntiles; t += t_stride, ++nt) {
      mbar_wait(&bars->q_ready, nt & 1);              // this tile's Q is in TMEM
      for (int iloc = 0; iloc < nblk; ++iloc, ++it) {
        const int i = it;                             // global block index: ring / buffer parity run across tiles
        const int s = i % STAGES, ph = (i / STAGES) & 1, j = i & 1;
        OG_TRACE_EVT(0, i);
        mbar_wait(&bars->k_full[s], ph);
        // S_i goes where P_lo of block i-2 was: P.V_{i-2} reads P_lo with its FIRST eight MMAs and signals plo_free, so this
        // QK^T does not wait for the other sixteen (the per-buffer chain QK -> softmax -> P.V -> QK paces the kernel)
        if (i >= 2) mbar_wait(&bars->plo_free[j], ((i - 2) >> 1) & 1);
        tc_fence_after();
        OG_TRACE_EVT(1, i);
        if (elect_one()) {
          const uint32_t khi = smem_u32(sK + s * k_stage_bytes<DH, CG>()), klo = khi + K_HALF;
          const uint32_t d_s = tmem + COL_SP + 128 * j + 64 * ((i >> 1) & 1);     // the halves of an S/P buffer swap roles per use
#pragma unroll
          for (int kk = 0; kk < DH / 8; ++kk) {
            const uint32_t off = (kk / 4) * KBLK + (kk % 4) * 32;
            const uint64_t dhi = make_sdesc_sw128(khi + off), dlo = make_sdesc_sw128(klo + off);
            mma(d_s, tmem + COL_QLO + kk * 8, dhi, idesc_qk, kk ? 1u : 0u);
            mma(d_s, tmem + COL_QHI + kk * 8, dlo, idesc_qk, 1u);
            mma(d_s, tmem + COL_QHI + kk * 8, dhi, idesc_qk, 1u);
          }
          commit(&bars->k_empty[s]);
          commit(&bars->s_full[j]);
          if (iloc == nblk - 1) commit(&bars->q_free);   // the tile's last QK^T: Q can be replaced once it retires
        }
        __syncwarp();
      }
      }
    } else {
      const int ntot = nblk * ((a.ntiles - t_first + t_stride - 1) / t_stride);   // all key blocks of all my tiles
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
          const uint32_t vhi = smem_u32(sV + s * v_stage_bytes<DH, CG>()), vlo = vhi + V_HALF;
          const uint32_t hh = (i >> 1) & 1;
          const uint32_t p_hi = tmem + COL_SP + 128 * j + 64 * hh, p_lo = tmem + COL_SP + 128 * j + 64 * (hh ^ 1);
          const uint32_t d_o = tmem + COL_O + 64 * j;
#pragma unroll
          for (int kk = 0; kk < BNK / 8; ++kk) {             // the P_lo readers first: their half is the next S accumulator
            const uint32_t off = (kk / 4) * VBLK + (kk % 4) * 32;
            mma(d_o, p_lo + kk * 8, make_sdesc_sw128(vhi + off), idesc_pv, kk ? 1u : 0u);
          }
          commit(&bars->plo_free[j]);
#pragma unroll
          for (int kk = 0; kk < BNK / 8; ++kk) {
            const uint32_t off = (kk / 4) * VBLK + (kk % 4) * 32;
            const uint64_t dhi = make_sdesc_sw128(vhi + off), dlo = make_sdesc_sw128(vlo + off);
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
    // Two warpgroups share every key block: warpgroup g owns logit columns [32g, 32g+32) of the 64-key block and
    // output channels [g*DH/2, (g+1)*DH/2).  Thread = one query row in both.  The row max of a block needs both
    // halves: one smem exchange + one 256-thread named barrier per block.  This halves the softmax latency per
    // block (the tensor pipe stalls whenever P_i is not ready within one QK + one PV of MMA time).
    constexpr int HD = DH / 2;                       // output channels per warpgroup
    const int g = warp >> 2;
    const int qd = warp & 3;
    const int trow = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float* xch = reinterpret_cast<float*>(bars + 1);                   // [2 parities][2 warpgroups][128] row maxima, then [2][128] sums
    const float c1 = a.scale * LOG2E;
    int it = 0, nt = 0;                              // key blocks / tiles done so far (buffer parities run across tiles)
    float4 qv[HD / 4];                               // this thread's half Q row of the NEXT tile to start
    constexpr bool STAGED = CG == 2;                 // row-coalesced global access through per-warp smem tiles (see stage_bytes)
    constexpr int LPR = HD / 4, RPI = 32 / LPR;      // lanes per row / rows per warp instruction when a row is HD floats
    float* qst = xch + 8 * 128 + warp * (2 * 32 * HD);   // this warp's Q staging tile [32][HD] (16-byte chunks XOR-swizzled by row)
    float* ost = qst + 32 * HD;                      //             output staging tile
    auto load_q = [&](const TilePos& p) {
      const int h = p.h, b = p.b;
      if constexpr (STAGED) {                        // cp.async: 8 (4) lanes fetch one row's 128 (64) bytes; lands during the tile
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
    auto write_q = [&]() {                           // qv -> split -> TMEM (A operand of every QK^T of a tile)
      if constexpr (STAGED) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) qv[c] = *reinterpret_cast<const float4*>(qst + lane * HD + ((c ^ (lane & (LPR - 1))) * 4));
        __syncwarp();                                // the tile may be refilled by the next load_q
      }
      if constexpr (HD == 32) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          split_tf32_fast(qv[c].x, hi[4 * c + 0], lo[4 * c + 0]); split_tf32_fast(qv[c].y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32_fast(qv[c].z, hi[4 * c + 2], lo[4 * c + 2]); split_tf32_fast(qv[c].w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        tmem_st_32x32(tmem + lane_base + COL_QHI + g * HD, hi);
        tmem_st_32x32(tmem + lane_base + COL_QLO + g * HD, lo);
      } else {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          split_tf32_fast(qv[c].x, hi[4 * c + 0], lo[4 * c + 0]); split_tf32_fast(qv[c].y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32_fast(qv[c].z, hi[4 * c + 2], lo[4 * c + 2]); split_tf32_fast(qv[c].w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        tmem_st_32x16(tmem + lane_base + COL_QHI + g * HD, hi);
        tmem_st_32x16(tmem + lane_base + COL_QLO + g * HD, lo);
      }
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
    tp = tile_next(tp);                              // from here on: the NEXT tile of this CTA
    if (nt == 0) write_q();                          // later tiles: written at the end of the previous tile (see below)
    if (t + t_stride < a.ntiles) load_q(tp);         // next tile's Q row: in flight during this whole tile

    float acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    // Online softmax in the exp2 domain.  c1 = scale*log2(e); mc = fl(m*c1) is the running max in that domain and
    // is used consistently for p = 2^(s*c1 - mc) (one FFMA: the product is exact inside the fma, only the small
    // difference is rounded) and for the block-to-block correction 2^(mc_old - mc_new).
    float m_run = -CUDART_INF_F, mc_run = -CUDART_INF_F, l_run = 0.f, corr_prev = 0.f;

    auto fold_o = [&](int i, float corr) {           // acc = acc * corr + my channels of O_i   (i = global block index)
      const int j = i & 1, jph = (i >> 1) & 1;
      mbar_wait(&bars->o_full[j], jph);
      tc_fence_after();
      const uint32_t op = tmem + lane_base + COL_O + 64 * j + g * HD;
      if constexpr (HD == 32) {
        uint32_t o[32];
        tmem_ld_32x32(op, o);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = fmaf(acc[c], corr, __uint_as_float(o[c]));
      } else {
        uint32_t o[16];
        tmem_ld_32x16(op, o);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = fmaf(acc[c], corr, __uint_as_float(o[c]));
      }
      tc_fence_before();
      arrive_leader(&bars->o_empty[j]);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(7, i); // O_i folded
    };

#pragma unroll 1
    for (int iloc = 0; iloc < nblk; ++iloc, ++it) {
      const int i = it;                              // global block index
      const int j = i & 1, jph = (i >> 1) & 1;
      const uint32_t hh = (i >> 1) & 1;              // which half of the buffer holds S_i / P_hi (the other one gets P_lo)
      const uint32_t sp = tmem + lane_base + COL_SP + 128 * j + 64 * hh + 32 * g;      // my 32 columns of S_i / P_hi
      const uint32_t sp_lo = tmem + lane_base + COL_SP + 128 * j + 64 * (hh ^ 1) + 32 * g;
      const int kbase = iloc * BNK + 32 * g;
      mbar_wait(&bars->s_full[j], jph);
      tc_fence_after();
      if (warp == 0 && lane == 0) OG_TRACE_EVT(5, i); // softmax: S_i observed
      if (iloc == nblk - 1 && t + t_stride < a.ntiles) {
        // The tile's last QK^T has retired (q_free completes together with this s_full): hand the NEXT tile's Q to the tensor
        // pipe now, before this block's softmax - its first QK^T then run under this tile's last softmax / P.V / epilogue
        // (written after the last P hand-over, the next tile's first S arrived ~2400 cycles later: event trace).
        mbar_wait(&bars->q_free, nt & 1);
        tc_fence_after();
        write_q();
      }
      uint32_t s[32], lo[32];
      tmem_ld_32x32(sp, s);
      tmem_wait_ld();
      if (kbase + 32 > a.nk) {                       // key tail (last block only): mask keys >= nk
#pragma unroll
        for (int c = 0; c < 32; ++c) if (kbase + c >= a.nk) s[c] = __float_as_uint(-CUDART_INF_F);
      }
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 32; c += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 1])));
      xch[(j * 2 + g) * 128 + trow] = mx;            // exchange the half-row maxima
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, xch[(j * 2 + (g ^ 1)) * 128 + trow]);
      const float m_new = fmaxf(m_run, mx);          // raw logits (scale > 0 commutes with max)
      const float mc = m_new * c1;
      const float corr = ex2_approx(mc_run - mc);    // 2^(-inf) = 0 on the first block
      float r0 = 0.f, r1 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), c1, -mc));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), c1, -mc));
        r0 += p0; r1 += p1;
        split_tf32_fast(p0, s[c], lo[c]);
        split_tf32_fast(p1, s[c + 1], lo[c + 1]);
      }
      tmem_st_32x32(sp, s);                          // P_hi over S, P_lo into the other half (where P_hi of block i-2 was:
      tmem_st_32x32(sp_lo, lo);                      // this thread has already waited for P.V_{i-2} in fold_o(i-2))
      tmem_wait_st();
      tc_fence_before();
      arrive_leader(&bars->p_full[j]);
      if (warp == 0 && lane == 0) OG_TRACE_EVT(6, i); // softmax: P_i handed over
      l_run = fmaf(l_run, corr, r0 + r1);            // partial row sum over my 32 columns (same max in both warpgroups)
      m_run = m_new; mc_run = mc;
      // P_i is on its way; now fold O_{i-1} (computed with m_{i-1}: its correction is the one saved last iteration).
      // Doing this AFTER the softmax keeps P_i off the critical path; O_{i-1}'s buffer is only needed again by PV_{i+1}.
      if (iloc >= 1) fold_o(i - 1, corr_prev);
      corr_prev = corr;
    }
    fold_o(it - 1, corr_prev);

    // total row sum = sum of the two warpgroups' partial sums (buffer alternates per tile: one barrier suffices)
    float* xl = xch + 4 * 128 + (nt & 1) * 256;
    xl[g * 128 + trow] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.f / (l_run + xl[(g ^ 1) * 128 + trow]);
    if constexpr (STAGED) {
#pragma unroll
      for (int c = 0; c < HD / 4; ++c)
        *reinterpret_cast<float4*>(ost + lane * HD + ((c ^ (lane & (LPR - 1))) * 4)) =
            make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
      __syncwarp();
      const int r_in = lane / LPR, ch = lane % LPR;
#pragma unroll
      for (int k = 0; k < LPR; ++k) {
        const int row = k * RPI + r_in, orow_g = grow - lane + row;          // global query row of staging row `row`
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
    tc_fence_before();
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();                 // the peer may still be reading my smem / signalling my barriers
  if (warp == 9) { tc_fence_after(); if (CG == 2) tmem_dealloc_pair<tca::TMEM_COLS>(tmem); else tmem_dealloc<tca::TMEM_COLS>(tmem); }
}

// khi/klo: [batch*nk, ldk];  vthi/vtlo: [batch*d, ldvt]
template <int DH, int CG>
inline int attention_tc_launch_t(const TcAttnArgs& a, const float* khi, const float* klo, int64_t ldk, const float* vthi,
                                 const float* vtlo, int64_t ldvt, cudaStream_t stream) {
  using namespace tca;
  CUtensorMap mkh, mkl, mvh, mvl;
  int rc;
  if ((rc = tc::make_tmap_2d(&mkh, khi, (uint64_t)a.batch * a.nk, (uint64_t)a.d, (uint64_t)ldk, BNK / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d(&mkl, klo, (uint64_t)a.batch * a.nk, (uint64_t)a.d, (uint64_t)ldk, BNK / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d(&mvh, vthi, (uint64_t)a.batch * a.d, (uint64_t)a.nk, (uint64_t)ldvt, DH / CG)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d(&mvl, vtlo, (uint64_t)a.batch * a.d, (uint64_t)a.nk, (uint64_t)ldvt, DH / CG)) != OG_OK) return rc;
  static DeviceFlags attr_set;
  if (attr_set.once()) {
    OG_CUDA(cudaFuncSetAttribute(attention_tc_kernel<DH, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<DH, CG>()));
  }
  TcAttnArgs ap = a;
  ap.nqg = cdiv(cdiv(a.nq, BM), CG);                                             // CG = 2: an odd last query block gets a phantom partner
  ap.ntiles = ap.nqg * a.num_heads * a.batch;
  const int sms = device_info().ok ? device_info().sm_count : 148;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(std::min(ap.ntiles, sms / CG) * CG);                        // persistent: one CTA (pair) per SM (pair)
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem_bytes<DH, CG>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = tc::pdl_mode() ? 2 : 1;
  OG_CUDA(cudaLaunchKernelEx(&cfg, attention_tc_kernel<DH, CG>, mkh, mkl, mvh, mvl, ap));
  launch_counter()++;
  return OG_OK;
}

// The cta_group::2 (CTA pair, M = 256) form is the default: each SM stages and reads only half of every K / V^T tile, which
// takes the kernel off the shared-memory-bandwidth limit of the single-CTA form (0.563 -> 0.538 ms per self layer at the
// headline shape).  OG_ATTN_PAIR=0 / og_set_tuning select the single-CTA form; both are parity-tested.
inline int& attention_tc_pair_mode() {
  static int v = [] { const char* e = getenv("OG_ATTN_PAIR"); return e ? atoi(e) : 1; }();
  return v;
}

inline int attention_tc_launch(const TcAttnArgs& a, const float* khi, const float* klo, int64_t ldk, const float* vthi,
                               const float* vtlo, int64_t ldvt, int head_dim, cudaStream_t stream) {
  const bool pair = attention_tc_pair_mode() != 0;
  if (head_dim == 64) return pair ? attention_tc_launch_t<64, 2>(a, khi, klo, ldk, vthi, vtlo, ldvt, stream)
                                  : attention_tc_launch_t<64, 1>(a, khi, klo, ldk, vthi, vtlo, ldvt, stream);
  if (head_dim == 32) return pair ? attention_tc_launch_t<32, 2>(a, khi, klo, ldk, vthi, vtlo, ldvt, stream)
                                  : attention_tc_launch_t<32, 1>(a, khi, klo, ldk, vthi, vtlo, ldvt, stream);
  return fail(OG_EUNSUPPORTED, "attention_tc: head_dim %d not in {32, 64}", head_dim);
}

inline bool attention_tc_eligible(int head_dim, int64_t ldq, int64_t ldk, int64_t ldvt, int64_t ldo) {
  return (head_dim == 32 || head_dim == 64) && ldq % 4 == 0 && ldk % 4 == 0 && ldvt % 4 == 0 && ldo % 4 == 0;
}

}  // namespace og
