// Persistent tcgen05 GEMM, 3xTF32 with CHUNKED accumulation (production linear-layer kernel).
//     Y = epi(alpha * [A|A2] . B^T + bias)
//
// Why chunked: the tensor core adds into its TMEM accumulator with truncation (measured: error grows
// ~linearly with the number of accumulation steps, 4e-6 relative at K=512, 10x worse than fp32 FMA).
// Here the accumulator only ever holds K=64 worth of products (24 MMAs); the epilogue warpgroups drain it
// into registers with round-to-nearest adds while the tensor pipe fills the other accumulator buffer.
// That restores fp32-GEMM accuracy (3-6e-7 relative) at no cost in tensor throughput.
//
// Data movement: one TMA producer thread streams, per 32-wide K block, the raw fp32 A tile (128 x 32) and
// the pre-split B_hi / B_lo tiles into a 128B-swizzled smem ring.  Four converter warps read their A row
// from smem (conflict-free through the swizzle), split it hi/lo in registers and tcgen05.st it into a TMEM
// ring, so A never occupies MMA-side smem bandwidth.  Row-major outputs leave through swizzled smem staging
// tiles and TMA stores.
// 16 warps: WG0 / WG1 = accumulate + epilogue for output columns [0,64) / [64,128) (64 fp32 accumulators per
// thread, so 128 registers per thread suffice for every role), WG2 = A converters, warp 12 = TMA producer,
// warp 13 = MMA issue + TMEM alloc.  One CTA per SM, static round-robin tile schedule (n fastest).
//
// PAIR = 1 (cta_group::2): two CTAs of a cluster work on 256 rows x 128 columns.  Each stages only HALF of the
// B tile (64 of its 128 rows); the leader CTA issues M = 256 MMAs that read B from both shared memories and A /
// D from both tensor memories.  Per CTA and K block this removes 16 KB of TMA writes and 24 KB of MMA operand
// reads from shared memory - the resource the single-CTA form is bound by (event trace: 128 KB of smem traffic
// per K block = 1000 cycles at 128 B/clk against 768 cycles of MMA work).
#pragma once
#include "tc_common.cuh"
#include "linear_tc.cuh"     // TcLinearArgs
#include <algorithm>
#include <stdlib.h>

namespace og {
namespace tcl2 {
constexpr int BM = 128, BN = 128, BK = 32;
constexpr int CHUNK_KB = 2;               // K blocks per accumulator chunk (K = 64)
constexpr int TILE_BYTES = 128 * BK * 4;  // 16 KB: a [128 x 32] fp32 tile
constexpr int THREADS = 512;
constexpr int TMEM_COLS = 512;            // acc buffers [0,128) [128,256); A ring 256 + 64 s (hi 32 | lo 32)
constexpr int COL_A = 256;
constexpr int OUT_TILE = 128 * 32 * 4;    // one staged [128 rows x 32 cols] output chunk (16 KB)
constexpr int OUT_BYTES = 2 * 2 * OUT_TILE;  // 2 epilogue warpgroups x 2 buffers
constexpr int MAX_STAGES = 4;

template <int PAIR> struct Cfg {
  static constexpr int STAGES = PAIR ? 4 : 3;                 // smem ring depth = TMEM A ring depth
  static constexpr int BROWS = BN / (PAIR ? 2 : 1);           // rows of B staged by one CTA
  static constexpr int B_TILE = BROWS * BK * 4;               // bytes of its B_hi (or B_lo) part
  static constexpr int STAGE_BYTES = TILE_BYTES + 2 * B_TILE; // A raw | B_hi | B_lo
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + OUT_BYTES + 2048;
};

struct __align__(16) Barriers {
  uint64_t a_land[MAX_STAGES], b_full[MAX_STAGES], empty[MAX_STAGES], a_full[MAX_STAGES], a_empty[MAX_STAGES];
  uint64_t acc_full[2], acc_empty[2], r_full[2];     // r_full[h]: residual chunks of warpgroup h landed in its staging buffers
  uint32_t tmem_base;
  alignas(16) float bias[2][BN];        // per-tile epilogue vectors (tile parity), staged by each epilogue warpgroup for its
  alignas(16) float rscale[2][BN];      // own 64 columns at tile start, while it waits for the first accumulator chunk
};

struct Sched { int ntmg, ntn, ngroups, nkb, nchunks; };   // ntmg: groups of (PAIR ? 2 : 1) m-tiles; ngroups = ntn * ntmg * batch
}  // namespace tcl2

// OUTK prunes the epilogue at compile time (the all-in-one epilogue was ~4000 SASS instructions of mutually exclusive paths;
// the event trace showed ~900 cycles per 32-column chunk for ~100 useful instructions - instruction fetch, not math):
//   0 generic (every combination, plain-store fallbacks)   1 Y only, through TMA (bias / ReLU; residual only with RTMA)
//   2 Yhi / Ylo only, through TMA                          3 Ythi / Ytlo only (transposed split outputs, direct stores)
template <int PAIR, int RTMA, int OUTK>
__global__ void __launch_bounds__(tcl2::THREADS, 1) linear_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                      const __grid_constant__ CUtensorMap map_a2,
                                                                      const __grid_constant__ CUtensorMap map_bhi,
                                                                      const __grid_constant__ CUtensorMap map_blo,
                                                                      const __grid_constant__ CUtensorMap map_y,
                                                                      const __grid_constant__ CUtensorMap map_yhi,
                                                                      const __grid_constant__ CUtensorMap map_ylo,
                                                                      const __grid_constant__ CUtensorMap map_r,
                                                                      TcLinearArgs a, tcl2::Sched sc, int y_tma) {
  constexpr bool r_tma = RTMA != 0;                          // residual tile arrives by TMA (separate instantiation: keeps the plain epilogue lean)
  using namespace tcl2;
  using namespace tc;
  using C = Cfg<PAIR>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_TILE = C::B_TILE, NC = PAIR ? 2 : 1;
  launch_dependents();                                       // the next kernel may take this SM as soon as this CTA leaves it
  extern __shared__ uint8_t og_tcl2_smem_raw[];
  uint8_t* smem = tc::align_smem_1024(og_tcl2_smem_raw);
  uint8_t* s_out = smem + STAGES * STAGE_BYTES;                          // [2 warpgroups][2 buffers][16 KB]
  Barriers* bars = reinterpret_cast<Barriers*>(s_out + OUT_BYTES);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;                // 0 = leader of the pair
  const int g_first = blockIdx.x / NC, g_stride = gridDim.x / NC;      // pair (or CTA) index / count

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bars->a_land[i], 1); mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->empty[i], 4 + 1);                     // 4 converter warps + the MMA commit
      mbar_init(&bars->a_full[i], 4 * NC); mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 8 * NC); mbar_init(&bars->r_full[i], 1); }
    fence_barrier_init();
    prefetch_tensormap(&map_a); prefetch_tensormap(&map_a2);
    prefetch_tensormap(&map_bhi); prefetch_tensormap(&map_blo);
  }
  if (PAIR) cluster_sync_all();                              // the peer's barriers exist before anyone signals them
  if (warp == 13) { if (PAIR) tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base); else tmem_alloc<TMEM_COLS>(&bars->tmem_base); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  grid_dependency_wait();                                    // nothing above touches memory written by the previous kernel
  // one arrival per warp (after every lane has completed and fenced its own TMEM / smem accesses) on a barrier that
  // lives in the leader CTA
  auto arrive_leader = [&](uint64_t* bar) {
    __syncwarp();
    if (lane == 0) { if (!PAIR || crank == 0) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); }
  };
  auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };

  // group tile t = NC consecutive m-tiles x one n-tile; this CTA takes m-tile number `crank` of the group
  auto tile_coords = [&](int t, int& m0, int& n0, int& bz) {
    n0 = (t % sc.ntn) * BN;
    m0 = (((t / sc.ntn) % sc.ntmg) * NC + (int)crank) * BM;
    bz = t / (sc.ntn * sc.ntmg);
  };

  if (warp >= 12) {
  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;                                              // global k-block counter (ring position)
      for (int t = g_first; t < sc.ngroups; t += g_stride) {
        int m0, n0, bz; tile_coords(t, m0, n0, bz);
        const int arow = bz * a.rows + m0;                     // batches are dense: row index into [batch*rows, K]
        const int brow = n0 + bz * a.b_rows_per_batch + (int)crank * C::BROWS;     // my part of the B tile's rows
        for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&bars->empty[s], ph ^ 1);                  // my converters and the MMAs released stage s
          OG_TRACE_EVT(0, it);
          uint8_t* dst = smem + s * STAGE_BYTES;
          const int k = kb * BK;
          mbar_arrive_expect_tx(&bars->a_land[s], TILE_BYTES);
          if (k < a.k1) tma_load_2d(dst, &map_a, &bars->a_land[s], k, arow);
          else          tma_load_2d(dst, &map_a2, &bars->a_land[s], k - a.k1, arow);
          if (crank == 0) mbar_arrive_expect_tx(&bars->b_full[s], NC * 2 * B_TILE);   // both CTAs' B parts report to the leader
          if (PAIR) {
            tma_load_2d_pair(dst + TILE_BYTES, &map_bhi, &bars->b_full[s], k, brow);
            tma_load_2d_pair(dst + TILE_BYTES + B_TILE, &map_blo, &bars->b_full[s], k, brow);
          } else {
            tma_load_2d(dst + TILE_BYTES, &map_bhi, &bars->b_full[s], k, brow);
            tma_load_2d(dst + TILE_BYTES + B_TILE, &map_blo, &bars->b_full[s], k, brow);
          }
        }
      }
    }
  } else if (warp == 13 && crank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only when paired)
    const uint32_t idesc = make_idesc_tf32(BM * NC, BN);
    int it = 0, g = 0;                                         // k-block and chunk counters
    for (int t = g_first; t < sc.ngroups; t += g_stride) {
      for (int c = 0; c < sc.nchunks; ++c, ++g) {
        const int buf = g & 1, gph = (g >> 1) & 1;
        mbar_wait(&bars->acc_empty[buf], gph ^ 1);
        tc_fence_after();
        const int kb_end = min((c + 1) * CHUNK_KB, sc.nkb);
        for (int kb = c * CHUNK_KB; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&bars->b_full[s], ph);                     // B tiles landed (in both CTAs)
          OG_TRACE_EVT(3, it);
          mbar_wait(&bars->a_full[s], ph);                     // A split written to TMEM (in both CTAs)
          tc_fence_after();
          OG_TRACE_EVT(4, it);
          if (elect_one()) {
            const uint32_t bhi = smem_u32(smem + s * STAGE_BYTES + TILE_BYTES), blo = bhi + B_TILE;
            const uint32_t d = tmem + buf * 128;
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {
              const uint64_t dbhi = make_sdesc_sw128(bhi + kk * 32), dblo = make_sdesc_sw128(blo + kk * 32);
              const uint32_t ahi = tmem + COL_A + s * 64 + kk * 8, alo = ahi + 32;
              const uint32_t acc0 = (kb > c * CHUNK_KB || kk) ? 1u : 0u;
              if (PAIR) {
                umma_tf32_ts_pair(d, alo, dbhi, idesc, acc0);
                umma_tf32_ts_pair(d, ahi, dblo, idesc, 1u);
                umma_tf32_ts_pair(d, ahi, dbhi, idesc, 1u);
              } else {
                umma_tf32_ts(d, alo, dbhi, idesc, acc0);
                umma_tf32_ts(d, ahi, dblo, idesc, 1u);
                umma_tf32_ts(d, ahi, dbhi, idesc, 1u);
              }
            }
            commit(&bars->empty[s]);
            commit(&bars->a_empty[s]);
            if (kb == kb_end - 1) commit(&bars->acc_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ A converters: smem fp32 -> split -> TMEM
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int t = g_first; t < sc.ngroups; t += g_stride) {
      for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&bars->a_land[s], ph);
        if (warp == 8 && lane == 0) OG_TRACE_EVT(1, it);
        const uint8_t* arow = smem + s * STAGE_BYTES + trow * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (trow & 7)) * 16));   // undo the 128B swizzle
          split_tf32_fast(v.x, hi[4 * c + 0], lo[4 * c + 0]); split_tf32_fast(v.y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32_fast(v.z, hi[4 * c + 2], lo[4 * c + 2]); split_tf32_fast(v.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->empty[s]);           // this warp is done with the smem A tile
        mbar_wait(&bars->a_empty[s], ph ^ 1);
        tc_fence_after();
        if (warp == 8 && lane == 0) OG_TRACE_EVT(7, it);
        const uint32_t taddr = tmem + lane_base + COL_A + s * 64;
        tmem_st_32x32(taddr, hi);
        tmem_st_32x32(taddr + 32, lo);
        tmem_wait_st();
        tc_fence_before();
        arrive_leader(&bars->a_full[s]);
        if (warp == 8 && lane == 0) OG_TRACE_EVT(2, it);
      }
    }
  } else {
    // ------------------------------------------------------------------ accumulate (RN, registers) + epilogue
    const int half = warp >> 2;                                // 0: output columns [0,64), 1: [64,128) of the tile
    constexpr int HN = BN / 2;
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int wg_tid = threadIdx.x & 127;
    const bool vec_r = a.R && (a.ldr % 4 == 0) && (a.strideR % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.R) & 15) == 0);
    int g = 0, ntile = 0;
    for (int t = g_first; t < sc.ngroups; t += g_stride, ++ntile) {
      int m0, n0, bz; tile_coords(t, m0, n0, bz);
      float racc[HN];
#pragma unroll
      for (int j = 0; j < HN; ++j) racc[j] = 0.f;
      const int pb = ntile & 1;
      {                                                        // global-load latency hides behind the main loop
        const int c = half * HN + (wg_tid & 63), col = n0 + c;
        if (wg_tid < 64) bars->bias[pb][c] = (a.bias && col < a.nout) ? __ldg(a.bias + col) : 0.f;
        else             bars->rscale[pb][c] = (a.rscale && col < a.nout) ? __ldg(a.rscale + col) : 1.f;
      }
      if (r_tma && wg_tid == 0) {
        // residual tile of this warpgroup's 64 columns -> its two staging buffers, in flight during the whole main loop
        // (per-thread row loads of R cost what the per-thread row stores did: ~1/3 of the tile time on the fc.3 GEMM)
        tma_store_wait_read<0>();                              // the previous tile's stores no longer read the buffers
        mbar_arrive_expect_tx(&bars->r_full[half], 2 * OUT_TILE);
        tma_load_3d(s_out + (half * 2 + 0) * OUT_TILE, &map_r, &bars->r_full[half], n0 + half * HN, m0, bz);
        tma_load_3d(s_out + (half * 2 + 1) * OUT_TILE, &map_r, &bars->r_full[half], n0 + half * HN + 32, m0, bz);
      }
      for (int c = 0; c < sc.nchunks; ++c, ++g) {
        const int buf = g & 1, gph = (g >> 1) & 1;
        mbar_wait(&bars->acc_full[buf], gph);
        tc_fence_after();
        if (warp == 0 && lane == 0) OG_TRACE_EVT(5, g);
#pragma unroll
        for (int ch = 0; ch < HN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem + lane_base + buf * 128 + half * HN + ch * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[ch * 32 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        arrive_leader(&bars->acc_empty[buf]);
        if (warp == 0 && lane == 0) OG_TRACE_EVT(6, g);
      }
      // ---- epilogue for this tile (overlaps the next tile's first chunks)
      if (warp == 0 && lane == 0) OG_TRACE_EVT(8, ntile);
      constexpr bool GEN = OUTK == 0;
      const bool has_y = GEN ? a.Y != nullptr : OUTK == 1, has_yhi = GEN ? a.Yhi != nullptr : OUTK == 2;
      const bool has_yt = GEN && a.Yt != nullptr, has_ythi = GEN ? a.Ythi != nullptr : OUTK == 3;
      const bool use_tma = GEN ? y_tma != 0 : OUTK != 3;
      const int grow = m0 + trow;
      const bool row_ok = grow < a.rows;
      const float* Rrow = (GEN && !r_tma && a.R) ? a.R + (int64_t)bz * a.strideR + (int64_t)grow * a.ldr : nullptr;
      const int64_t yoff = (int64_t)bz * a.strideY + (int64_t)grow * a.ldy;
      const int64_t ytoff = (int64_t)bz * a.strideYt + grow;
      // One warpgroup barrier per tile makes the staged bias visible and tells everyone that the previous tile's TMA
      // stores have finished reading the two staging buffers (they were issued a whole main loop ago).
      if (wg_tid == 0 && !r_tma) tma_store_wait_read<0>();
      asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
      if (warp == 0 && lane == 0) OG_TRACE_EVT(9, ntile);
      // staged stores are issued in groups of up to two buffers: fill, one fence + barrier, then the TMA stores
      int npend = 0;
      const CUtensorMap* pmap0 = nullptr; const CUtensorMap* pmap1 = nullptr; int pcol0 = 0, pcol1 = 0;
      auto flush = [&]() {
        if (npend == 0) return;
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
        if (wg_tid == 0) {
          tma_store_3d(pmap0, s_out + (half * 2 + 0) * OUT_TILE, pcol0, m0, bz);
          if (npend == 2) tma_store_3d(pmap1, s_out + (half * 2 + 1) * OUT_TILE, pcol1, m0, bz);
          tma_store_commit();
        }
        npend = 0;
      };
#pragma unroll
      for (int cc = 0; cc < HN / 32; ++cc) {                   // 32-column chunks of this warpgroup's half
        const int cl = half * HN + cc * 32;                    // column inside the tile
        const int cb = n0 + cl;
        if (cb >= a.nout) continue;                            // uniform across the warpgroup
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&bars->bias[pb][cl + j]);
          y[j] = fmaf(racc[cc * 32 + j], a.alpha, bv.x);         y[j + 1] = fmaf(racc[cc * 32 + j + 1], a.alpha, bv.y);
          y[j + 2] = fmaf(racc[cc * 32 + j + 2], a.alpha, bv.z); y[j + 3] = fmaf(racc[cc * 32 + j + 3], a.alpha, bv.w);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (r_tma) {                                           // residual chunk cc sits (swizzled) in staging buffer cc
          if (cc == 0) mbar_wait(&bars->r_full[half], ntile & 1);   // ... and the result goes back out through the same buffer
          const uint8_t* rsrc = s_out + (half * 2 + cc) * OUT_TILE + trow * 128;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 r = *reinterpret_cast<const float4*>(rsrc + ((c4 ^ (trow & 7)) * 16));
            const float4 sv = *reinterpret_cast<const float4*>(&bars->rscale[pb][cl + 4 * c4]);
            y[4 * c4] = fmaf(sv.x, r.x, y[4 * c4]);         y[4 * c4 + 1] = fmaf(sv.y, r.y, y[4 * c4 + 1]);
            y[4 * c4 + 2] = fmaf(sv.z, r.z, y[4 * c4 + 2]); y[4 * c4 + 3] = fmaf(sv.w, r.w, y[4 * c4 + 3]);
          }
        } else if (Rrow && row_ok) {
          if (cb + 31 < a.nout && vec_r) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r = *reinterpret_cast<const float4*>(Rrow + cb + j);
              const float4 sv = *reinterpret_cast<const float4*>(&bars->rscale[pb][cl + j]);
              y[j] = fmaf(sv.x, r.x, y[j]); y[j + 1] = fmaf(sv.y, r.y, y[j + 1]);
              y[j + 2] = fmaf(sv.z, r.z, y[j + 2]); y[j + 3] = fmaf(sv.w, r.w, y[j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (cb + j < a.nout) y[j] = fmaf(bars->rscale[pb][cl + j], Rrow[cb + j], y[j]);
          }
        }
        // transposed outputs: for a fixed column the 32 lanes write 32 consecutive rows (already coalesced)
        if (has_yt && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (cb + j < a.nout) a.Yt[ytoff + (int64_t)(cb + j) * a.ldyt] = y[j];
        }
        // row-major outputs: the chunk is staged in 128B-swizzled smem and written by one TMA store (full 128-byte
        // segments per row, clipped to the tensor bounds); measured before: per-thread row stores took ~10K of the
        // ~16K cycles per tile.
        auto stage_store = [&](const CUtensorMap* map, const uint32_t (&v)[32]) {
          if (npend == 2) {                                    // both buffers hold unsent data: send, then wait until they were read
            flush();
            if (wg_tid == 0) tma_store_wait_read<0>();
            asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
          }
          if (warp == 0 && lane == 0) OG_TRACE_EVT(10 + cc, ntile);
          uint8_t* dst = s_out + (half * 2 + npend) * OUT_TILE + trow * 128;     // r_tma: npend == cc, the residual chunk's own buffer
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(dst + ((c ^ (trow & 7)) * 16)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          if (npend == 0) { pmap0 = map; pcol0 = cb; } else { pmap1 = map; pcol1 = cb; }
          ++npend;
          if (warp == 0 && lane == 0) OG_TRACE_EVT(12 + cc, ntile);
        };
        if (use_tma) {
          if (has_y) {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(y[j]);
            stage_store(&map_y, v);
          }
          if (has_yhi) {
            uint32_t yh[32], yl[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) split_tf32(y[j], yh[j], yl[j]);
            stage_store(&map_yhi, yh);
            stage_store(&map_ylo, yl);
          }
        } else if (GEN && row_ok) {                            // unaligned row pitch: plain stores
#pragma unroll
          for (int j = 0; j < 32; ++j) if (cb + j < a.nout) {
            if (has_y) a.Y[yoff + cb + j] = y[j];
            if (has_yhi) { uint32_t h, l; split_tf32(y[j], h, l); a.Yhi[yoff + cb + j] = __uint_as_float(h); a.Ylo[yoff + cb + j] = __uint_as_float(l); }
          }
        }
        if (has_ythi && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (cb + j < a.nout) {
            uint32_t h, l; split_tf32(y[j], h, l);
            const int64_t o = ytoff + (int64_t)(cb + j) * a.ldyt;
            a.Ythi[o] = __uint_as_float(h); a.Ytlo[o] = __uint_as_float(l);
          }
        }
      }
      flush();
      if (warp == 0 && lane == 0) OG_TRACE_EVT(14, ntile);
    }
    if (wg_tid == 0) tma_store_wait_all<0>();                  // smem must outlive the last store's reads
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                              // no CTA leaves while the peer may still read its smem / signal its barriers
  if (warp == 13) { tc_fence_after(); if (PAIR) tmem_dealloc_pair<tcl2::TMEM_COLS>(tmem); else tmem_dealloc<tcl2::TMEM_COLS>(tmem); }
}

inline bool linear_tc2_eligible(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb) {
  if (!linear_tc_eligible(a, Bhi, Blo, ldb)) return false;
  if (a.batch > 1 && a.strideA != (int64_t)a.rows * a.lda) return false;            // dense batches (one 2-D tensor map)
  if (a.A2 && (a.k1 % 32 != 0 || (a.batch > 1 && a.strideA2 != (int64_t)a.rows * a.lda2))) return false;
  return true;
}

template <int PAIR>
inline int linear_tc2_launch_t(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb, int64_t b_total_rows,
                               cudaStream_t stream) {
  using namespace tcl2;
  using C = Cfg<PAIR>;
  constexpr int NC = PAIR ? 2 : 1;
  const int K = a.k1 + a.k2;
  CUtensorMap ma, ma2, mhi, mlo;
  int rc;
  const uint64_t arows = (uint64_t)a.batch * a.rows;
  if ((rc = tc::make_tmap_2d(&ma, a.A, arows, (uint64_t)a.k1, (uint64_t)a.lda, BM)) != OG_OK) return rc;
  if (a.A2) { if ((rc = tc::make_tmap_2d(&ma2, a.A2, arows, (uint64_t)a.k2, (uint64_t)a.lda2, BM)) != OG_OK) return rc; }
  else ma2 = ma;
  if ((rc = tc::make_tmap_2d(&mhi, Bhi, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, C::BROWS)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d(&mlo, Blo, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, C::BROWS)) != OG_OK) return rc;
  // row-major outputs go out through TMA stores when their rows are 16-byte aligned
  CUtensorMap my = ma, myh = ma, myl = ma;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int y_tma = (a.Y || a.Yhi) && a.ldy % 4 == 0 && a.strideY % 4 == 0 && (!a.Y || al16(a.Y)) && (!a.Yhi || (al16(a.Yhi) && al16(a.Ylo)));
  if (y_tma) {
    if (a.Y && (rc = tc::make_tmap_3d(&my, a.Y, a.batch, a.rows, a.nout, a.ldy, a.strideY, BM)) != OG_OK) return rc;
    if (a.Yhi && (rc = tc::make_tmap_3d(&myh, a.Yhi, a.batch, a.rows, a.nout, a.ldy, a.strideY, BM)) != OG_OK) return rc;
    if (a.Yhi && (rc = tc::make_tmap_3d(&myl, a.Ylo, a.batch, a.rows, a.nout, a.ldy, a.strideY, BM)) != OG_OK) return rc;
  }
  // residual through TMA as well: needs the staging buffers for itself, so only without split outputs; the specialised
  // instantiation writes Y only (a transposed output - the context descriptors of image 0 - takes the generic one)
  CUtensorMap mr = ma;
  const int r_tma = y_tma && a.R && a.Y && !a.Yhi && !a.Yt && !a.Ythi && a.ldr % 4 == 0 && a.strideR % 4 == 0 && al16(a.R) && a.nout % 32 == 0;
  if (r_tma && (rc = tc::make_tmap_3d(&mr, a.R, a.batch, a.rows, a.nout, a.ldr, a.strideR, BM)) != OG_OK) return rc;
  // epilogue specialisation (see the kernel's OUTK): everything else takes the generic instantiation
  int outk = 0;
  if (!a.Yt && (r_tma || !a.R)) {
    if (y_tma && a.Y && !a.Yhi && !a.Ythi) outk = 1;
    else if (y_tma && !a.Y && a.Yhi && !a.Ythi) outk = 2;
    else if (!a.Y && !a.Yhi && a.Ythi) outk = 3;
  }
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                        TcLinearArgs, Sched, int);
  static const Kern kerns[5] = {linear_tc2_kernel<PAIR, 0, 0>, linear_tc2_kernel<PAIR, 0, 1>, linear_tc2_kernel<PAIR, 0, 2>,
                                linear_tc2_kernel<PAIR, 0, 3>, linear_tc2_kernel<PAIR, 1, 1>};
  static DeviceFlags attr_set;
  if (attr_set.once()) {
    for (Kern k : kerns) OG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  const Kern kern = r_tma ? kerns[4] : kerns[outk];
  Sched sc;
  sc.ntmg = cdiv(cdiv(a.rows, BM), NC); sc.ntn = cdiv(a.nout, BN); sc.ngroups = sc.ntmg * sc.ntn * a.batch;
  sc.nkb = cdiv(K, BK); sc.nchunks = cdiv(sc.nkb, CHUNK_KB);
  const int sms = device_info().ok ? device_info().sm_count : 148;
  const int nclusters = std::min(sc.ngroups, sms / NC);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * NC);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = tc::pdl_mode() ? 2 : 1;
  OG_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, ma2, mhi, mlo, my, myh, myl, mr, a, sc, y_tma));
  launch_counter()++;
  return OG_OK;
}

// The cta_group::2 form is the default (OG_GEMM_PAIR=0 / og_set_tuning select the single-CTA form; both parity-tested).
// Event traces: single CTA ~1400 cycles per K block (128 KB of shared-memory traffic each), pair ~1050 (88 KB) against 768 of
// MMA work.  An earlier measurement had the pair SLOWER (137 vs 186 TF/s): the peer CTA's hand-offs used
// mbarrier.arrive.release.cluster, which cost ~1000+ cycles each; with the default-semantics remote arrive they are ~free.
inline int& linear_tc2_pair_mode() {
  static int v = [] { const char* e = getenv("OG_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  return v;
}

inline int linear_tc2_launch(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb, int64_t b_total_rows,
                             cudaStream_t stream) {
  const bool pair = linear_tc2_pair_mode() != 0 && cdiv(a.rows, tcl2::BM) >= 2;     // tiny problems: no phantom m-tiles
  return pair ? linear_tc2_launch_t<1>(a, Bhi, Blo, ldb, b_total_rows, stream)
              : linear_tc2_launch_t<0>(a, Bhi, Blo, ldb, b_total_rows, stream);
}

}  // namespace og
