// tcgen05 GEMM with fp32-grade accuracy from tf32 tensor cores ("3xTF32"):
//     Y = epi(alpha * [A|A2] . B^T + bias),   A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi   (fp32 accumulate in TMEM)
//
// * B (weights, or an activation tensor that a producer kernel already wrote split) lives in HBM as two
//   tf32-exact tensors B_hi, B_lo; 128 x 32-float tiles are staged by TMA into 128B-swizzled shared memory.
// * A (fp32 activations) never touches shared memory in the default mode: the four converter warps read
//   their row straight from global memory, split it in registers and tcgen05.st it into TMEM, from where
//   the MMA consumes it (A-from-TMEM form).  MODE_SS keeps A in shared memory instead (written by the same
//   warps in the swizzled K-major layout) - kept as a cross-check of the TMEM-operand path.
// * one elected thread issues tcgen05.mma (M=128, N=128, K=8 per instruction); smem / TMEM stages are
//   recycled through mbarriers signalled by tcgen05.commit; the same four warps run the epilogue
//   (tcgen05.ld -> bias / ReLU / residual -> fp32, split and/or transposed stores).
// One CTA per 128 x 128 output tile; two CTAs fit per SM (256 TMEM columns, 96 KB smem each) so one CTA's
// epilogue overlaps the other's main loop.
#pragma once
#include "tc_common.cuh"

namespace og {

struct TcLinearArgs {
  const float* A;  int64_t lda;  int64_t strideA;
  const float* A2; int64_t lda2; int64_t strideA2;
  int k1, k2;
  int b_rows_per_batch;                 // B tile row offset per batch item (0: B shared by the batch)
  const float* bias;
  int rows, nout, batch;
  float alpha;
  int relu;
  const float* R; int64_t ldr; int64_t strideR;
  const float* rscale;
  float* Y;   float* Yhi;  float* Ylo;  int64_t ldy;  int64_t strideY;     // row-major outputs (any may be null)
  float* Yt;  float* Ythi; float* Ytlo; int64_t ldyt; int64_t strideYt;    // transposed outputs [nout, rows]
};

namespace tcl {
constexpr int BM = 128, BN = 128, BK = 32;
constexpr int B_STAGES = 3, A_STAGES = 2;
constexpr int TILE_BYTES = BN * BK * 4;                  // 16 KB: one hi or lo tile
constexpr int TMEM_COLS = 256;                           // D: [0,128)  A stages: 128 + 64*s (hi 32 | lo 32)
constexpr int THREADS = 192;
enum { MODE_TS = 0, MODE_SS = 1 };

struct __align__(8) Barriers {
  uint64_t b_full[B_STAGES], b_empty[B_STAGES], a_full[A_STAGES], a_empty[A_STAGES], d_full;
  uint32_t tmem_base;
};
constexpr int smem_bytes(int mode) {
  return 1024 /*align slack*/ + B_STAGES * 2 * TILE_BYTES + (mode == MODE_SS ? A_STAGES * 2 * TILE_BYTES : 0) + 256;
}
}  // namespace tcl

template <int MODE>
__global__ void __launch_bounds__(tcl::THREADS) linear_tc_kernel(const __grid_constant__ CUtensorMap map_bhi,
                                                                 const __grid_constant__ CUtensorMap map_blo,
                                                                 TcLinearArgs a) {
  using namespace tcl;
  using namespace tc;
  extern __shared__ uint8_t og_tcl_smem_raw[];
  uint8_t* smem = tc::align_smem_1024(og_tcl_smem_raw);
  uint8_t* sB = smem;                                                     // [B_STAGES][hi|lo][16 KB]
  uint8_t* sA = smem + B_STAGES * 2 * TILE_BYTES;                         // MODE_SS only
  Barriers* bars = reinterpret_cast<Barriers*>(smem + B_STAGES * 2 * TILE_BYTES + (MODE == MODE_SS ? A_STAGES * 2 * TILE_BYTES : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, bz = blockIdx.z;
  const int K = a.k1 + a.k2;
  const int nkb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], 1); }
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&bars->a_full[i], 128); mbar_init(&bars->a_empty[i], 1); }
    mbar_init(&bars->d_full, 1);
    fence_barrier_init();
    prefetch_tensormap(&map_bhi);
    prefetch_tensormap(&map_blo);
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      const int brow = n0 + bz * a.b_rows_per_batch;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % B_STAGES, ph = (kb / B_STAGES) & 1;
        mbar_wait(&bars->b_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bars->b_full[s], 2 * TILE_BYTES);
        tma_load_2d(sB + (s * 2 + 0) * TILE_BYTES, &map_bhi, &bars->b_full[s], kb * BK, brow);
        tma_load_2d(sB + (s * 2 + 1) * TILE_BYTES, &map_blo, &bars->b_full[s], kb * BK, brow);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_tf32(BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int bs = kb % B_STAGES, bph = (kb / B_STAGES) & 1;
      const int as = kb % A_STAGES, aph = (kb / A_STAGES) & 1;
      mbar_wait(&bars->b_full[bs], bph);
      mbar_wait(&bars->a_full[as], aph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t bhi = smem_u32(sB + (bs * 2 + 0) * TILE_BYTES), blo = smem_u32(sB + (bs * 2 + 1) * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t dbhi = make_sdesc_sw128(bhi + kk * 32), dblo = make_sdesc_sw128(blo + kk * 32);
          const uint32_t acc0 = (kb | kk) ? 1u : 0u;
          if (MODE == MODE_TS) {
            const uint32_t ahi = tmem + 128 + as * 64 + kk * 8, alo = ahi + 32;
            umma_tf32_ts(tmem, alo, dbhi, idesc, acc0);
            umma_tf32_ts(tmem, ahi, dblo, idesc, 1u);
            umma_tf32_ts(tmem, ahi, dbhi, idesc, 1u);
          } else {
            const uint32_t sahi = smem_u32(sA + (as * 2 + 0) * TILE_BYTES), salo = smem_u32(sA + (as * 2 + 1) * TILE_BYTES);
            const uint64_t dahi = make_sdesc_sw128(sahi + kk * 32), dalo = make_sdesc_sw128(salo + kk * 32);
            umma_tf32_ss(tmem, dalo, dbhi, idesc, acc0);
            umma_tf32_ss(tmem, dahi, dblo, idesc, 1u);
            umma_tf32_ss(tmem, dahi, dbhi, idesc, 1u);
          }
        }
        umma_commit(&bars->b_empty[bs]);
        umma_commit(&bars->a_empty[as]);
        if (kb == nkb - 1) umma_commit(&bars->d_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ converter warps, then epilogue
    const int q = warp & 3;                                  // TMEM lane quarter this warp may touch
    const int trow = q * 32 + lane;                          // tile row owned by this thread
    const int grow = m0 + trow;
    const bool row_ok = grow < a.rows;
    const float* Arow = a.A + (int64_t)bz * a.strideA + (int64_t)grow * a.lda;
    const float* A2row = a.A2 ? a.A2 + (int64_t)bz * a.strideA2 + (int64_t)grow * a.lda2 : nullptr;

    auto load_kb = [&](int kb, float4 (&v)[8]) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k = kb * BK + c * 4;
        if (row_ok && k < K) {
          const float* src = (k < a.k1) ? (Arow + k) : (A2row + (k - a.k1));
          v[c] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    float4 cur[8], nxt[8];
    load_kb(0, cur);
    for (int kb = 0; kb < nkb; ++kb) {
      if (kb + 1 < nkb) load_kb(kb + 1, nxt);
      const int as = kb % A_STAGES, aph = (kb / A_STAGES) & 1;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        split_tf32(cur[c].x, hi[4 * c + 0], lo[4 * c + 0]);
        split_tf32(cur[c].y, hi[4 * c + 1], lo[4 * c + 1]);
        split_tf32(cur[c].z, hi[4 * c + 2], lo[4 * c + 2]);
        split_tf32(cur[c].w, hi[4 * c + 3], lo[4 * c + 3]);
      }
      mbar_wait(&bars->a_empty[as], aph ^ 1);
      if (MODE == MODE_TS) {
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 128 + as * 64;
        tmem_st_32x32(taddr, hi);
        tmem_st_32x32(taddr + 32, lo);
        tmem_wait_st();
        tc_fence_before();
      } else {
        uint8_t* thi = sA + (as * 2 + 0) * TILE_BYTES + trow * 128;
        uint8_t* tlo = sA + (as * 2 + 1) * TILE_BYTES + trow * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int pc = (c ^ (trow & 7)) * 16;             // 128B swizzle: 16-byte chunk index XOR (row mod 8)
          *reinterpret_cast<uint4*>(thi + pc) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          *reinterpret_cast<uint4*>(tlo + pc) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
        fence_proxy_async();
      }
      mbar_arrive(&bars->a_full[as]);
#pragma unroll
      for (int c = 0; c < 8; ++c) cur[c] = nxt[c];
    }

    // ---- epilogue
    mbar_wait(&bars->d_full, 0);
    tc_fence_after();
    const float* Rrow = a.R ? a.R + (int64_t)bz * a.strideR + (int64_t)grow * a.ldr : nullptr;
    const int64_t yoff = (int64_t)bz * a.strideY + (int64_t)grow * a.ldy;
    const int64_t ytoff = (int64_t)bz * a.strideYt + grow;
    const bool vec_ok = (a.ldy % 4 == 0) && (a.strideY % 4 == 0);
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + ch * 32, acc);
      tmem_wait_ld();
      const int cb = n0 + ch * 32;
      if (cb >= a.nout) break;
      float y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = cb + j;
        float v = __uint_as_float(acc[j]) * a.alpha;
        if (c < a.nout) {
          if (a.bias) v += __ldg(a.bias + c);
          if (a.relu) v = fmaxf(v, 0.f);
          if (Rrow && row_ok) { const float rv = Rrow[c]; v = a.rscale ? fmaf(__ldg(a.rscale + c), rv, v) : (v + rv); }
        }
        y[j] = v;
      }
      const bool need_split = a.Yhi || a.Ythi;
      uint32_t yh[32], yl[32];
      if (need_split) {
#pragma unroll
        for (int j = 0; j < 32; ++j) split_tf32(y[j], yh[j], yl[j]);
      }
      if (row_ok) {
        if (vec_ok && cb + 31 < a.nout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (a.Y) *reinterpret_cast<float4*>(a.Y + yoff + cb + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
            if (a.Yhi) {
              *reinterpret_cast<uint4*>(a.Yhi + yoff + cb + j) = make_uint4(yh[j], yh[j + 1], yh[j + 2], yh[j + 3]);
              *reinterpret_cast<uint4*>(a.Ylo + yoff + cb + j) = make_uint4(yl[j], yl[j + 1], yl[j + 2], yl[j + 3]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cb + j < a.nout) {
              if (a.Y) a.Y[yoff + cb + j] = y[j];
              if (a.Yhi) { a.Yhi[yoff + cb + j] = __uint_as_float(yh[j]); a.Ylo[yoff + cb + j] = __uint_as_float(yl[j]); }
            }
          }
        }
        // transposed outputs: for a fixed column the 32 lanes write 32 consecutive rows (coalesced)
        if (a.Yt || a.Ythi) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cb + j < a.nout) {
              const int64_t o = ytoff + (int64_t)(cb + j) * a.ldyt;
              if (a.Yt) a.Yt[o] = y[j];
              if (a.Ythi) { a.Ythi[o] = __uint_as_float(yh[j]); a.Ytlo[o] = __uint_as_float(yl[j]); }
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<tcl::TMEM_COLS>(tmem); }
}

// Bhi/Blo: [b_total_rows, K] row-major fp32 (tf32-exact values), row stride ldb.
template <int MODE>
inline int linear_tc_launch_mode(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb,
                                 int64_t b_total_rows, cudaStream_t stream) {
  using namespace tcl;
  const int K = a.k1 + a.k2;
  CUtensorMap mhi, mlo;
  int rc = tc::make_tmap_2d(&mhi, Bhi, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, BN);
  if (rc != OG_OK) return rc;
  rc = tc::make_tmap_2d(&mlo, Blo, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, BN);
  if (rc != OG_OK) return rc;
  static DeviceFlags attr_set;
  if (attr_set.once()) {
    OG_CUDA(cudaFuncSetAttribute(linear_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(MODE)));
  }
  dim3 grid(cdiv(a.nout, BN), cdiv(a.rows, BM), a.batch);
  linear_tc_kernel<MODE><<<grid, THREADS, smem_bytes(MODE), stream>>>(mhi, mlo, a);
  OG_LAUNCH_CHECK("linear_tc_kernel");
  launch_counter()++;
  return OG_OK;
}

inline bool linear_tc_eligible(const TcLinearArgs& a, const float* Bhi, const float* Blo, int64_t ldb) {
  const int K = a.k1 + a.k2;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return K >= 32 && a.k1 % 4 == 0 && a.k2 % 4 == 0 && a.lda % 4 == 0 && a.strideA % 4 == 0 && al16(a.A) &&
         (!a.A2 || (a.lda2 % 4 == 0 && a.strideA2 % 4 == 0 && al16(a.A2))) && ldb % 4 == 0 && al16(Bhi) && al16(Blo);
}

// elementwise x -> (hi, lo) tf32 split of a flat buffer (weights at pack time)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi,
                                                          float* __restrict__ lo, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t h, l;
    tc::split_tf32(src[i], h, l);
    hi[i] = __uint_as_float(h); lo[i] = __uint_as_float(l);
  }
}

}  // namespace og
