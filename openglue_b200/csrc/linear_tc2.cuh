// Persistent tcgen05 GEMM, 3xTF32 with CHUNKED accumulation (production linear-layer kernel).
//     Y = epi(alpha * [A|A2] . B^T + bias)
//
// Why chunked: the tensor core adds into its TMEM accumulator with truncation (measured: error grows
// ~linearly with the number of accumulation steps, 4e-6 relative at K=512, 10x worse than fp32 FMA).
// Here the accumulator only ever holds K=64 worth of products (24 MMAs); a second warpgroup drains it
// into registers with round-to-nearest adds while the tensor pipe fills the other accumulator buffer.
// That restores fp32-GEMM accuracy (3-6e-7 relative) at no cost in tensor throughput.
//
// Data movement: one TMA producer thread streams, per 32-wide K block, the raw fp32 A tile (128 x 32)
// and the pre-split B_hi / B_lo tiles (128 x 32 each) into a 4-deep 128B-swizzled smem ring.  The four
// converter warps read their A row from smem (conflict-free through the swizzle), split it hi/lo in
// registers and tcgen05.st it into a 4-deep TMEM ring, so A never occupies MMA-side smem bandwidth.
// 16 warps: WG0 / WG1 = accumulate + epilogue for output columns [0,64) / [64,128) (64 fp32 accumulators per
// thread, so 128 registers per thread suffice for every role), WG2 = A converters, warp 12 = TMA producer,
// warp 13 = MMA issue + TMEM alloc.
// One CTA per SM, static round-robin tile schedule (n fastest, so CTAs that run together share A in L2).
// Thread-block clusters of CL CTAs along M share the B tile: every CTA TMA-loads 1/CL of its rows with
// .multicast::cluster into all CL shared memories, cutting B's L2->SM traffic by CL (measured before: the
// kernel moved 6-7.5 TB/s out of L2, i.e. it was L2-bound with B = 2/3 of the bytes).  Stage recycling across
// CTAs: a producer tells its peers when its own stage is free and issues only when all peers said the same.
#pragma once
#include "tc_common.cuh"
#include "linear_tc.cuh"     // TcLinearArgs
#include <algorithm>
#include <stdlib.h>

namespace og {
namespace tcl2 {
constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;                 // smem ring (A raw + B hi + B lo); the TMEM A ring has the same depth
constexpr int CHUNK_KB = 2;               // K blocks per accumulator chunk (K = 64)
constexpr int TILE_BYTES = 128 * BK * 4;  // 16 KB
constexpr int STAGE_BYTES = 3 * TILE_BYTES;
constexpr int THREADS = 512;
constexpr int TMEM_COLS = 512;            // acc buffers [0,128) [128,256); A ring 256 + 64 s (hi 32 | lo 32), s < STAGES
constexpr int COL_A = 256;

struct __align__(16) Barriers {
  uint64_t full[STAGES], empty[STAGES], a_full[STAGES], a_empty[STAGES], acc_full[2], acc_empty[2], peer_free[STAGES];
  uint32_t tmem_base;
  alignas(16) float bias[BN];           // per-tile epilogue vectors staged by the epilogue warps (read as float4)
  alignas(16) float rscale[BN];
};
constexpr int OUT_TILE = 128 * 32 * 4;    // one staged [128 rows x 32 cols] output chunk (16 KB)
constexpr int OUT_BYTES = 2 * 2 * OUT_TILE;  // 2 epilogue warpgroups x 2 buffers
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + OUT_BYTES + 2048;

struct Sched { int ntmg, ntn, ngroups, nkb, nchunks; };   // ntmg: groups of CL m-tiles; ngroups = ntn * ntmg * batch
}  // namespace tcl2

template <int CL>
__global__ void __launch_bounds__(tcl2::THREADS, 1) linear_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                      const __grid_constant__ CUtensorMap map_a2,
                                                                      const __grid_constant__ CUtensorMap map_bhi,
                                                                      const __grid_constant__ CUtensorMap map_blo,
                                                                      const __grid_constant__ CUtensorMap map_y,
                                                                      const __grid_constant__ CUtensorMap map_yhi,
                                                                      const __grid_constant__ CUtensorMap map_ylo,
                                                                      TcLinearArgs a, tcl2::Sched sc, int y_tma) {
  using namespace tcl2;
  using namespace tc;
  extern __shared__ uint8_t og_tcl2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(og_tcl2_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_out = smem + STAGES * STAGE_BYTES;                          // [2 warpgroups][2 buffers][16 KB]
  Barriers* bars = reinterpret_cast<Barriers*>(s_out + OUT_BYTES);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform by construction (setmaxnreg needs it)
  const int lane = threadIdx.x & 31;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  const int g_first = blockIdx.x / CL, g_stride = gridDim.x / CL;     // cluster index / number of clusters

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 129);       // 128 converter threads + MMA commit
      mbar_init(&bars->a_full[i], 128); mbar_init(&bars->a_empty[i], 1);
      mbar_init(&bars->peer_free[i], CL > 1 ? CL - 1 : 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 256); }
    fence_barrier_init();
    prefetch_tensormap(&map_a); prefetch_tensormap(&map_a2);
    prefetch_tensormap(&map_bhi); prefetch_tensormap(&map_blo);
  }
  if (warp == 13) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                            // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // group tile t = CL consecutive m-tiles x one n-tile; this CTA takes m-tile number `crank` of the group
  auto tile_coords = [&](int t, int& m0, int& n0, int& bz) {
    n0 = (t % sc.ntn) * BN;
    m0 = (((t / sc.ntn) % sc.ntmg) * CL + (int)crank) * BM;
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
        const int brow = n0 + bz * a.b_rows_per_batch;
        for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&bars->empty[s], ph ^ 1);                  // my consumers released stage s
          OG_TRACE_EVT(0, it);
          if (CL > 1) {
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t)CL; ++r) if (r != crank) mbar_arrive_remote(&bars->peer_free[s], r);
            mbar_wait(&bars->peer_free[s], ph);                // ... and so did every peer's (they will receive my B slice)
          }
          mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
          uint8_t* dst = smem + s * STAGE_BYTES;
          const int k = kb * BK;
          if (k < a.k1) tma_load_2d(dst, &map_a, &bars->full[s], k, arow);
          else          tma_load_2d(dst, &map_a2, &bars->full[s], k - a.k1, arow);
          if (CL > 1) {                                        // my 128/CL rows of B_hi / B_lo go to every CTA of the cluster
            constexpr int SL = BN / CL;
            tma_load_2d_mcast(dst + TILE_BYTES + crank * SL * 128, &map_bhi, &bars->full[s], k, brow + crank * SL, (uint16_t)((1u << CL) - 1));
            tma_load_2d_mcast(dst + 2 * TILE_BYTES + crank * SL * 128, &map_blo, &bars->full[s], k, brow + crank * SL, (uint16_t)((1u << CL) - 1));
          } else {
            tma_load_2d(dst + TILE_BYTES, &map_bhi, &bars->full[s], k, brow);
            tma_load_2d(dst + 2 * TILE_BYTES, &map_blo, &bars->full[s], k, brow);
          }
        }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_tf32(BM, BN);
    int it = 0, g = 0;                                         // k-block and chunk counters
    for (int t = g_first; t < sc.ngroups; t += g_stride) {
      for (int c = 0; c < sc.nchunks; ++c, ++g) {
        const int buf = g & 1, gph = (g >> 1) & 1;
        mbar_wait(&bars->acc_empty[buf], gph ^ 1);
        tc_fence_after();
        const int kb_end = min((c + 1) * CHUNK_KB, sc.nkb);
        for (int kb = c * CHUNK_KB; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&bars->full[s], ph);                       // B tiles landed
          OG_TRACE_EVT(3, it);
          mbar_wait(&bars->a_full[s], ph);                     // A split written to TMEM
          tc_fence_after();
          OG_TRACE_EVT(4, it);
          if (elect_one()) {
            const uint32_t bhi = smem_u32(smem + s * STAGE_BYTES + TILE_BYTES), blo = bhi + TILE_BYTES;
            const uint32_t d = tmem + buf * 128;
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {
              const uint64_t dbhi = make_sdesc_sw128(bhi + kk * 32), dblo = make_sdesc_sw128(blo + kk * 32);
              const uint32_t ahi = tmem + COL_A + s * 64 + kk * 8, alo = ahi + 32;
              umma_tf32_ts(d, alo, dbhi, idesc, (kb > c * CHUNK_KB || kk) ? 1u : 0u);
              umma_tf32_ts(d, ahi, dblo, idesc, 1u);
              umma_tf32_ts(d, ahi, dbhi, idesc, 1u);
            }
            umma_commit(&bars->empty[s]);
            umma_commit(&bars->a_empty[s]);
            if (kb == kb_end - 1) umma_commit(&bars->acc_full[buf]);
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
        mbar_wait(&bars->full[s], ph);
        if (warp == 8 && lane == 0) OG_TRACE_EVT(1, it);
        const uint8_t* arow = smem + s * STAGE_BYTES + trow * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (trow & 7)) * 16));   // undo the 128B swizzle
          split_tf32_fast(v.x, hi[4 * c + 0], lo[4 * c + 0]); split_tf32_fast(v.y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32_fast(v.z, hi[4 * c + 2], lo[4 * c + 2]); split_tf32_fast(v.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        mbar_arrive(&bars->empty[s]);                          // this thread is done with the smem A tile
        mbar_wait(&bars->a_empty[s], ph ^ 1);
        tc_fence_after();
        const uint32_t taddr = tmem + lane_base + COL_A + s * 64;
        tmem_st_32x32(taddr, hi);
        tmem_st_32x32(taddr + 32, lo);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->a_full[s]);
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
    int nstore = 0;                                            // staged stores issued by this warpgroup so far
    const bool vec_r = a.R && (a.ldr % 4 == 0) && (a.strideR % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.R) & 15) == 0);
    int g = 0;
    for (int t = g_first; t < sc.ngroups; t += g_stride) {
      int m0, n0, bz; tile_coords(t, m0, n0, bz);
      float racc[HN];
#pragma unroll
      for (int j = 0; j < HN; ++j) racc[j] = 0.f;
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
        mbar_arrive(&bars->acc_empty[buf]);
        if (warp == 0 && lane == 0) OG_TRACE_EVT(6, g);
      }
      // ---- epilogue for this tile (overlaps the next tile's first chunks)
      const int grow = m0 + trow;
      const bool row_ok = grow < a.rows;
      const float* Rrow = a.R ? a.R + (int64_t)bz * a.strideR + (int64_t)grow * a.ldr : nullptr;
      const int64_t yoff = (int64_t)bz * a.strideY + (int64_t)grow * a.ldy;
      const int64_t ytoff = (int64_t)bz * a.strideYt + grow;
      // stage this tile's bias / residual scale once (one element per thread) instead of 128 loads per thread
      asm volatile("bar.sync 1, 256;" ::: "memory");            // previous tile's readers (both halves) are done
      if (half == 0) {
        bars->bias[trow] = (a.bias && n0 + trow < a.nout) ? __ldg(a.bias + n0 + trow) : 0.f;
        bars->rscale[trow] = (a.rscale && n0 + trow < a.nout) ? __ldg(a.rscale + n0 + trow) : 1.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < HN / 32; ++cc) {                   // 32-column chunks of this warpgroup's half
        const int cl = half * HN + cc * 32;                    // column inside the tile
        const int cb = n0 + cl;
        if (cb >= a.nout) continue;                            // uniform across the warpgroup
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&bars->bias[cl + j]);
          y[j] = fmaf(racc[cc * 32 + j], a.alpha, bv.x);         y[j + 1] = fmaf(racc[cc * 32 + j + 1], a.alpha, bv.y);
          y[j + 2] = fmaf(racc[cc * 32 + j + 2], a.alpha, bv.z); y[j + 3] = fmaf(racc[cc * 32 + j + 3], a.alpha, bv.w);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (Rrow && row_ok) {
          if (cb + 31 < a.nout && vec_r) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r = *reinterpret_cast<const float4*>(Rrow + cb + j);
              const float4 sv = *reinterpret_cast<const float4*>(&bars->rscale[cl + j]);
              y[j] = fmaf(sv.x, r.x, y[j]); y[j + 1] = fmaf(sv.y, r.y, y[j + 1]);
              y[j + 2] = fmaf(sv.z, r.z, y[j + 2]); y[j + 3] = fmaf(sv.w, r.w, y[j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (cb + j < a.nout) y[j] = fmaf(bars->rscale[cl + j], Rrow[cb + j], y[j]);
          }
        }
        // transposed outputs: for a fixed column the 32 lanes write 32 consecutive rows (already coalesced)
        if (a.Yt && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (cb + j < a.nout) a.Yt[ytoff + (int64_t)(cb + j) * a.ldyt] = y[j];
        }
        // row-major outputs: the chunk is staged in 128B-swizzled smem and written by one TMA store (full 128-byte
        // segments per row, clipped to the tensor bounds); measured before: per-thread row stores took ~10K of the
        // ~16K cycles per tile.
        auto stage_store = [&](const CUtensorMap* map, const uint32_t (&v)[32]) {
          uint8_t* buf = s_out + (half * 2 + (nstore & 1)) * OUT_TILE;
          if (wg_tid == 0) tma_store_wait_read<1>();           // the store that last read this buffer is done with it
          asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
          uint8_t* dst = buf + trow * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(dst + ((c ^ (trow & 7)) * 16)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
          if (wg_tid == 0) { tma_store_3d(map, buf, cb, m0, bz); tma_store_commit(); }
          ++nstore;
        };
        if (y_tma) {
          if (a.Y) {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(y[j]);
            stage_store(&map_y, v);
          }
          if (a.Yhi) {
            uint32_t yh[32], yl[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) split_tf32(y[j], yh[j], yl[j]);
            stage_store(&map_yhi, yh);
            stage_store(&map_ylo, yl);
          }
        } else if (row_ok) {                                   // unaligned row pitch: plain stores
#pragma unroll
          for (int j = 0; j < 32; ++j) if (cb + j < a.nout) {
            if (a.Y) a.Y[yoff + cb + j] = y[j];
            if (a.Yhi) { uint32_t h, l; split_tf32(y[j], h, l); a.Yhi[yoff + cb + j] = __uint_as_float(h); a.Ylo[yoff + cb + j] = __uint_as_float(l); }
          }
        }
        if (a.Ythi && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (cb + j < a.nout) {
            uint32_t h, l; split_tf32(y[j], h, l);
            const int64_t o = ytoff + (int64_t)(cb + j) * a.ldyt;
            a.Ythi[o] = __uint_as_float(h); a.Ytlo[o] = __uint_as_float(l);
          }
        }
      }
    }
    if (wg_tid == 0) tma_store_wait_all<0>();                  // smem must outlive the last store's reads
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                            // no CTA leaves while a peer may still write its smem / barriers
  if (warp == 13) { tc_fence_after(); tmem_dealloc<tcl2::TMEM_COLS>(tmem); }
}

inline bool linear_tc2_eligible(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb) {
  if (!linear_tc_eligible(a, Bhi, Blo, ldb)) return false;
  if (a.batch > 1 && a.strideA != (int64_t)a.rows * a.lda) return false;            // dense batches (one 2-D tensor map)
  if (a.A2 && (a.k1 % 32 != 0 || (a.batch > 1 && a.strideA2 != (int64_t)a.rows * a.lda2))) return false;
  return true;
}

template <int CL>
inline int linear_tc2_launch_cl(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb, int64_t b_total_rows,
                                cudaStream_t stream) {
  using namespace tcl2;
  const int K = a.k1 + a.k2;
  CUtensorMap ma, ma2, mhi, mlo;
  int rc;
  const uint64_t arows = (uint64_t)a.batch * a.rows;
  if ((rc = tc::make_tmap_2d(&ma, a.A, arows, (uint64_t)a.k1, (uint64_t)a.lda, BM)) != OG_OK) return rc;
  if (a.A2) { if ((rc = tc::make_tmap_2d(&ma2, a.A2, arows, (uint64_t)a.k2, (uint64_t)a.lda2, BM)) != OG_OK) return rc; }
  else ma2 = ma;
  if ((rc = tc::make_tmap_2d(&mhi, Bhi, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, BN / CL)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d(&mlo, Blo, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, BN / CL)) != OG_OK) return rc;
  // row-major outputs go out through TMA stores when their rows are 16-byte aligned
  CUtensorMap my = ma, myh = ma, myl = ma;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int y_tma = (a.Y || a.Yhi) && a.ldy % 4 == 0 && a.strideY % 4 == 0 && (!a.Y || al16(a.Y)) && (!a.Yhi || (al16(a.Yhi) && al16(a.Ylo)));
  if (y_tma) {
    if (a.Y && (rc = tc::make_tmap_3d(&my, a.Y, a.batch, a.rows, a.nout, a.ldy, a.strideY, BM)) != OG_OK) return rc;
    if (a.Yhi && (rc = tc::make_tmap_3d(&myh, a.Yhi, a.batch, a.rows, a.nout, a.ldy, a.strideY, BM)) != OG_OK) return rc;
    if (a.Yhi && (rc = tc::make_tmap_3d(&myl, a.Ylo, a.batch, a.rows, a.nout, a.ldy, a.strideY, BM)) != OG_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    OG_CUDA(cudaFuncSetAttribute(linear_tc2_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Sched sc;
  sc.ntmg = cdiv(cdiv(a.rows, BM), CL); sc.ntn = cdiv(a.nout, BN); sc.ngroups = sc.ntmg * sc.ntn * a.batch;
  sc.nkb = cdiv(K, BK); sc.nchunks = cdiv(sc.nkb, CHUNK_KB);
  const int sms = device_info().ok ? device_info().sm_count : 148;
  const int nclusters = std::min(sc.ngroups, sms / CL);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CL);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  OG_CUDA(cudaLaunchKernelEx(&cfg, linear_tc2_kernel<CL>, ma, ma2, mhi, mlo, my, myh, myl, a, sc, y_tma));
  launch_counter()++;
  return OG_OK;
}

// Cluster size along M (B-tile multicast).  Measured on B200 (profiles/README.md): 2-CTA clusters are ~5% slower and
// 4-CTA clusters 2.8x slower than no clusters - the kernel is bound by the depth of its smem ring (bytes in flight per
// SM), which multicast does not change, and the cross-CTA stage handshake adds latency.  Default 1; OG_TC_CLUSTER=2|4
// keeps the path testable.
inline int linear_tc2_cluster_size() {
  static int cl = [] {
    const char* e = getenv("OG_TC_CLUSTER");
    int v = e ? atoi(e) : 1;
    return (v == 1 || v == 2 || v == 4) ? v : 1;
  }();
  return cl;
}

inline int linear_tc2_launch(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb, int64_t b_total_rows,
                             cudaStream_t stream) {
  int cl = linear_tc2_cluster_size();
  while (cl > 1 && cdiv(a.rows, tcl2::BM) < cl) cl >>= 1;              // tiny problems: no point in phantom m-tiles
  switch (cl) {
    case 4: return linear_tc2_launch_cl<4>(a, Bhi, Blo, ldb, b_total_rows, stream);
    case 2: return linear_tc2_launch_cl<2>(a, Bhi, Blo, ldb, b_total_rows, stream);
    default: return linear_tc2_launch_cl<1>(a, Bhi, Blo, ldb, b_total_rows, stream);
  }
}

}  // namespace og
