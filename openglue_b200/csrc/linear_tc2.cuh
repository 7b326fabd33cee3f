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
// Warp roles: 0 = TMA, 1 = MMA issue (+TMEM alloc), 2-5 = A converters, 6-9 = accumulate + epilogue.
// One CTA per SM, static round-robin tile schedule (n fastest, so CTAs that run together share A in L2).
#pragma once
#include "tc_common.cuh"
#include "linear_tc.cuh"     // TcLinearArgs

namespace og {
namespace tcl2 {
constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 4;                 // smem ring (A raw + B hi + B lo) and TMEM A ring
constexpr int CHUNK_KB = 2;               // K blocks per accumulator chunk (K = 64)
constexpr int TILE_BYTES = 128 * BK * 4;  // 16 KB
constexpr int STAGE_BYTES = 3 * TILE_BYTES;
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;            // acc buffers [0,128) [128,256); A ring 256 + 64 s (hi 32 | lo 32)
constexpr int COL_A = 256;

struct __align__(8) Barriers {
  uint64_t full[STAGES], empty[STAGES], a_full[STAGES], a_empty[STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 512;

struct Sched { int ntm, ntn, ntiles, nkb, nchunks; };
}  // namespace tcl2

__global__ void __maxnreg__(200) linear_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                      const __grid_constant__ CUtensorMap map_a2,
                                                                      const __grid_constant__ CUtensorMap map_bhi,
                                                                      const __grid_constant__ CUtensorMap map_blo,
                                                                      TcLinearArgs a, tcl2::Sched sc) {
  using namespace tcl2;
  using namespace tc;
  extern __shared__ uint8_t og_tcl2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(og_tcl2_smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 129);       // 128 converter threads + MMA commit
      mbar_init(&bars->a_full[i], 128); mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 128); }
    fence_barrier_init();
    prefetch_tensormap(&map_a); prefetch_tensormap(&map_a2);
    prefetch_tensormap(&map_bhi); prefetch_tensormap(&map_blo);
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  auto tile_coords = [&](int t, int& m0, int& n0, int& bz) {
    n0 = (t % sc.ntn) * BN;
    m0 = ((t / sc.ntn) % sc.ntm) * BM;
    bz = t / (sc.ntn * sc.ntm);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;                                              // global k-block counter (ring position)
      for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x) {
        int m0, n0, bz; tile_coords(t, m0, n0, bz);
        const int arow = bz * a.rows + m0;                     // batches are dense: row index into [batch*rows, K]
        const int brow = n0 + bz * a.b_rows_per_batch;
        for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&bars->empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
          uint8_t* dst = smem + s * STAGE_BYTES;
          const int k = kb * BK;
          if (k < a.k1) tma_load_2d(dst, &map_a, &bars->full[s], k, arow);
          else          tma_load_2d(dst, &map_a2, &bars->full[s], k - a.k1, arow);
          tma_load_2d(dst + TILE_BYTES, &map_bhi, &bars->full[s], k, brow);
          tma_load_2d(dst + 2 * TILE_BYTES, &map_blo, &bars->full[s], k, brow);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_tf32(BM, BN);
    int it = 0, g = 0;                                         // k-block and chunk counters
    for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x) {
      for (int c = 0; c < sc.nchunks; ++c, ++g) {
        const int buf = g & 1, gph = (g >> 1) & 1;
        mbar_wait(&bars->acc_empty[buf], gph ^ 1);
        tc_fence_after();
        const int kb_end = min((c + 1) * CHUNK_KB, sc.nkb);
        for (int kb = c * CHUNK_KB; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&bars->full[s], ph);                       // B tiles landed
          mbar_wait(&bars->a_full[s], ph);                     // A split written to TMEM
          tc_fence_after();
          if (lane == 0) {
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
  } else if (warp < 6) {
    // ------------------------------------------------------------------ A converters: smem fp32 -> split -> TMEM
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x) {
      for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&bars->full[s], ph);
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
      }
    }
  } else {
    // ------------------------------------------------------------------ accumulate (RN, registers) + epilogue
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool vec_ok = (a.ldy % 4 == 0) && (a.strideY % 4 == 0);
    const bool vec_r = a.R && (a.ldr % 4 == 0) && (a.strideR % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.R) & 15) == 0);
    int g = 0;
    for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x) {
      int m0, n0, bz; tile_coords(t, m0, n0, bz);
      float racc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) racc[j] = 0.f;
      for (int c = 0; c < sc.nchunks; ++c, ++g) {
        const int buf = g & 1, gph = (g >> 1) & 1;
        mbar_wait(&bars->acc_full[buf], gph);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem + lane_base + buf * 128 + ch * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[ch * 32 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        mbar_arrive(&bars->acc_empty[buf]);
      }
      // ---- epilogue for this tile (overlaps the next tile's first chunks)
      const int grow = m0 + trow;
      const bool row_ok = grow < a.rows;
      const float* Rrow = a.R ? a.R + (int64_t)bz * a.strideR + (int64_t)grow * a.ldr : nullptr;
      const int64_t yoff = (int64_t)bz * a.strideY + (int64_t)grow * a.ldy;
      const int64_t ytoff = (int64_t)bz * a.strideYt + grow;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {                   // not unrolled: keeps the kernel's code footprint small
        const int cb = n0 + ch * 32;
        if (cb >= a.nout || !row_ok) continue;
        const bool full = cb + 31 < a.nout;
        float y[32];
        switch (ch) {                                          // racc must stay statically indexed (registers)
#define OG_COPY_CHUNK(CH_) case CH_: { _Pragma("unroll") for (int j = 0; j < 32; ++j) y[j] = racc[CH_ * 32 + j] * a.alpha; } break;
          OG_COPY_CHUNK(0) OG_COPY_CHUNK(1) OG_COPY_CHUNK(2) default: OG_COPY_CHUNK(3)
#undef OG_COPY_CHUNK
        }
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (full || cb + j < a.nout) y[j] += __ldg(a.bias + cb + j);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (Rrow) {
          if (full && vec_r) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r = *reinterpret_cast<const float4*>(Rrow + cb + j);
              if (a.rscale) {
                y[j] = fmaf(__ldg(a.rscale + cb + j), r.x, y[j]); y[j + 1] = fmaf(__ldg(a.rscale + cb + j + 1), r.y, y[j + 1]);
                y[j + 2] = fmaf(__ldg(a.rscale + cb + j + 2), r.z, y[j + 2]); y[j + 3] = fmaf(__ldg(a.rscale + cb + j + 3), r.w, y[j + 3]);
              } else { y[j] += r.x; y[j + 1] += r.y; y[j + 2] += r.z; y[j + 3] += r.w; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (cb + j < a.nout) {
              const float rv = Rrow[cb + j];
              y[j] = a.rscale ? fmaf(__ldg(a.rscale + cb + j), rv, y[j]) : (y[j] + rv);
            }
          }
        }
        if (a.Y) {
          if (full && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(a.Y + yoff + cb + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (cb + j < a.nout) a.Y[yoff + cb + j] = y[j];
          }
        }
        if (a.Yt) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (full || cb + j < a.nout) a.Yt[ytoff + (int64_t)(cb + j) * a.ldyt] = y[j];
        }
        if (a.Yhi || a.Ythi) {
          uint32_t yh[32], yl[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) split_tf32(y[j], yh[j], yl[j]);
          if (a.Yhi) {
            if (full && vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                *reinterpret_cast<uint4*>(a.Yhi + yoff + cb + j) = make_uint4(yh[j], yh[j + 1], yh[j + 2], yh[j + 3]);
                *reinterpret_cast<uint4*>(a.Ylo + yoff + cb + j) = make_uint4(yl[j], yl[j + 1], yl[j + 2], yl[j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (cb + j < a.nout) {
                a.Yhi[yoff + cb + j] = __uint_as_float(yh[j]); a.Ylo[yoff + cb + j] = __uint_as_float(yl[j]);
              }
            }
          }
          if (a.Ythi) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (full || cb + j < a.nout) {
              const int64_t o = ytoff + (int64_t)(cb + j) * a.ldyt;
              a.Ythi[o] = __uint_as_float(yh[j]); a.Ytlo[o] = __uint_as_float(yl[j]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<tcl2::TMEM_COLS>(tmem); }
}

inline bool linear_tc2_eligible(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb) {
  if (!linear_tc_eligible(a, Bhi, Blo, ldb)) return false;
  if (a.batch > 1 && a.strideA != (int64_t)a.rows * a.lda) return false;            // dense batches (one 2-D tensor map)
  if (a.A2 && (a.k1 % 32 != 0 || (a.batch > 1 && a.strideA2 != (int64_t)a.rows * a.lda2))) return false;
  return true;
}

inline int linear_tc2_launch(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb, int64_t b_total_rows,
                             cudaStream_t stream) {
  using namespace tcl2;
  const int K = a.k1 + a.k2;
  CUtensorMap ma, ma2, mhi, mlo;
  int rc;
  const uint64_t arows = (uint64_t)a.batch * a.rows;
  if ((rc = tc::make_tmap_2d(&ma, a.A, arows, (uint64_t)a.k1, (uint64_t)a.lda, BM)) != OG_OK) return rc;
  if (a.A2) { if ((rc = tc::make_tmap_2d(&ma2, a.A2, arows, (uint64_t)a.k2, (uint64_t)a.lda2, BM)) != OG_OK) return rc; }
  else ma2 = ma;
  if ((rc = tc::make_tmap_2d(&mhi, Bhi, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, BN)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d(&mlo, Blo, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, BN)) != OG_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    OG_CUDA(cudaFuncSetAttribute(linear_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Sched sc;
  sc.ntm = cdiv(a.rows, BM); sc.ntn = cdiv(a.nout, BN); sc.ntiles = sc.ntm * sc.ntn * a.batch;
  sc.nkb = cdiv(K, BK); sc.nchunks = cdiv(sc.nkb, CHUNK_KB);
  const int sms = device_info().ok ? device_info().sm_count : 148;
  const int grid = sc.ntiles < sms ? sc.ntiles : sms;
  linear_tc2_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(ma, ma2, mhi, mlo, a, sc);
  OG_LAUNCH_CHECK("linear_tc2_kernel");
  launch_counter()++;
  return OG_OK;
}

}  // namespace og
