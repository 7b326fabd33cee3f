// Persistent tcgen05 GEMM with fp16 hi/lo operands ("3xFP16", kind::f16) and chunked accumulation:
//     Y = epi(alpha * [A|A2] . B^T + bias)
// Same numerics contract as csrc/linear_tc2.cuh (three products lo.hi + hi.lo + hi.hi, the TMEM accumulator only ever
// holds K = 64 and is drained into registers with round-to-nearest adds), at TWICE the MMA rate (K = 16 per instruction) and
// with HALF the B-operand bytes in shared memory - the resource the tf32 form is bound by.
//
// Range management (see tc_common.cuh): A is scaled on the fly by sA = f16_scale_for(max of the producers' tracked amax
// slots), B (weights) is pre-split with its own power-of-two scale sW (w_meta[0]); alpha / (sA sW) undoes both exactly.
// fp16 OUTPUTS (K, V^T for the attention kernel) need their scale BEFORE the first element is written, so it comes from a
// bound, not from the data: |y| <= |alpha| (amax(A) max_n ||W_n||_1 + max|b|)  (w_meta[1], w_meta[2], computed at pack time);
// the scale is published through scale_out.  fp32 outputs track their true amax into amax_out for the next consumer.
//
// Data movement per 64-wide K block: the raw fp32 A tile (two 128 x 32 TMA boxes, 32 KB) and the B_hi / B_lo fp16 tiles
// ([128 or 64 rows] x 64 halves = 128-byte swizzled rows).  Four converter warps read their A row from smem, scale, split
// and pack it (element 2c in the low half of 32-bit column c) and tcgen05.st it into a TMEM ring: A is a TMEM operand.
// 16 warps: WG0 / WG1 = accumulate + epilogue (output columns [0,64) / [64,128)), WG2 = converters, warps 12 / 14 = TMA producers
// of the A / B rings, warp 13 = MMA issue + TMEM alloc.  PAIR = cta_group::2 (256 rows x 128 columns per CTA pair, each CTA stages half of B).
//
// OUTK: 1 = fp32 Y through TMA (bias / ReLU, residual through TMA with RTMA, amax tracking)
//       2 = row-major fp16 hi / lo through TMA (K operand of the attention kernel)
//       3 = transposed fp16 hi / lo, direct stores (V^T operand)
//       4 = SEVERAL projections of the same A in one launch (Q | K | V or K | V): B = the stacked weights, the output kind of a
//           tile follows from its column block (kind = kind0 + n0 / kind_cols: 0 -> as OUTK 1, 1 -> as OUTK 2, 2 -> as OUTK 3), every
//           kind with its own weight scale / norm bound (w_meta + 4 per kind).  Same tiles, same arithmetic as the separate launches
//           (results are bit-identical); one launch instead of three, and 2 - 3 x the tiles per launch (shorter tail).
#pragma once
#include "tc_common.cuh"
#include "linear_tc2.cuh"   // linear_tc2_pair_mode (og_set_tuning / OG_GEMM_PAIR is shared by both forms)
#include <algorithm>
#include <stdlib.h>

// OG_GEMM_CONV2 (default 0: parity-clean on B200, 855.9 against 862.9 pairs/s for the whole step - no gain): the A converters load the whole fp32 row, hand the shared-memory slot back
// and only then split it, with packed fp32x2 arithmetic.
#ifndef OG_GEMM_CONV2
#define OG_GEMM_CONV2 0
#endif

namespace og {

struct F16LinearArgs {
  const float* A;  int64_t lda;  int64_t strideA;
  const float* A2; int64_t lda2; int64_t strideA2;
  int k1, k2;
  int b_rows_per_batch;                 // B tile row offset per batch item (0: B shared by the batch)
  const float* bias;
  int rows, nout, batch;
  float alpha;
  int relu;
  const float* R; int64_t ldr; int64_t strideR;          // fp32 residual (OUTK 1, through TMA)
  float* Y; int64_t ldy; int64_t strideY;                // OUTK 1
  __half* Yh; __half* Yl;                                // OUTK 2 (same ldy / strideY, in elements)
  __half* Yth; __half* Ytl; int64_t ldyt; int64_t strideYt;   // OUTK 3: [nout, rows] per batch item
  const float* amax_in[3];              // device scalars bounding |A| and |A2| (null entries ignored; at least one required)
  const float* w_meta;                  // device {scale of the pre-split B, max_n ||B_n||_1, max |bias|}
  float* amax_out;                      // optional: max |Y| (atomicMax; zeroed by the caller)
  float* scale_out;                     // OUTK 2 / 3: receives the scale the fp16 outputs were written with
  int swap_halves;                      // debug: pack A with element 2c in the HIGH half (probe of the TMEM operand layout)
  int nkinds, kind0, kind_cols;         // OUTK 4 (nkinds > 1): output kinds kind0 .. kind0 + nkinds - 1, kind_cols columns each (multiple of 128)
  float* scale_out_v;                   // OUTK 4: scale of the transposed (kind 2) output; scale_out is the row-major (kind 1) one
};

namespace tcf {
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int A_SUB = 128 * 32 * 4;       // 16 KB: one [128 x 32] fp32 TMA box
constexpr int A_BYTES = 2 * A_SUB;
constexpr int THREADS = 512;
constexpr int TMEM_COLS = 512;            // acc buffers [0,128) [128,256); A ring 256 + 64 s (hi 32 | lo 32 packed columns)
constexpr int COL_A = 256;
constexpr int OUT_TILE = 128 * 128;       // one staging tile: [128 rows x 32 fp32] or [128 rows x 64 fp16] (128-byte rows)
constexpr int OUT_BYTES = 2 * 2 * OUT_TILE;
constexpr int MAX_STAGES = 4;
constexpr int CHUNK_KB = 2;               // K blocks per accumulator chunk: K = 128 = 8 k-steps x 3 MMAs = 24 accumulations into TMEM,
                                          // the same count (and therefore the same truncation error) as the tf32 form's K = 64 chunks

// Two rings: the fp32 A tile's slot is free again as soon as the converters have read it (~1000 cycles after it landed), the B
// tiles' slot only when the MMAs that read it have retired (~2000 cycles later).  With ONE ring of three 48 KB stages a stage
// lived ~3300 cycles (load 1400 + convert 1000 + MMA 900: event trace profiles/r02_trace_gemm_f16_fc2_v1.txt) = one K block per
// 1100 cycles against 768 of MMA work; the 16 KB B stages are cheap to deepen.
template <int PAIR> struct Cfg {
  static constexpr int A_STAGES = 3;                          // = depth of the TMEM A ring
  static constexpr int B_STAGES = PAIR ? 4 : 2;
  static constexpr int BROWS = BN / (PAIR ? 2 : 1);
  static constexpr int B_TILE = BROWS * 128;                  // bytes of its B_hi (or B_lo) part: BROWS rows x 64 halves
  static constexpr int B_STAGE = 2 * B_TILE;
  static constexpr int SMEM_BYTES = 1024 + A_STAGES * A_BYTES + B_STAGES * B_STAGE + OUT_BYTES + 1536;   // 231936 of 232448 bytes
};

__device__ __forceinline__ unsigned long long pk2(float x, float y) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
// hi = fp16(x), lo = fp16(x - hi) for a packed pair of fp32 values (the scalar form is tc::split_f16x2)
__device__ __forceinline__ void split_f16x2_packed(unsigned long long x2, uint32_t& hi, uint32_t& lo) {
  float x0, x1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x2));
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 f = __half22float2(h);
  unsigned long long r2;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r2) : "l"(x2), "l"(pk2(f.x, f.y)));
  float r0, r1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r2));
  const __half2 l = __floats2half2_rn(r0, r1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct __align__(16) Barriers {
  uint64_t a_land[MAX_STAGES], a_free[MAX_STAGES], b_full[MAX_STAGES], b_free[MAX_STAGES], a_full[MAX_STAGES], a_empty[MAX_STAGES];
  uint64_t acc_full[2], acc_empty[2], r_full[2];
  uint32_t tmem_base;
  alignas(16) float bias[2][BN];
};
static_assert(sizeof(Barriers) <= 1536, "Barriers must fit the tail of the shared-memory budget");
struct Sched { int ntmg, ntn, ngroups, nkb, nchunks; };
}  // namespace tcf

template <int PAIR, int RTMA, int OUTK>
__global__ void __launch_bounds__(tcf::THREADS, 1) linear_f16_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_a2,
                                                                     const __grid_constant__ CUtensorMap map_bhi,
                                                                     const __grid_constant__ CUtensorMap map_blo,
                                                                     const __grid_constant__ CUtensorMap map_y,
                                                                     const __grid_constant__ CUtensorMap map_yh,
                                                                     const __grid_constant__ CUtensorMap map_yl,
                                                                     const __grid_constant__ CUtensorMap map_r,
                                                                     F16LinearArgs a, tcf::Sched sc) {
  using namespace tcf;
  using namespace tc;
  using C = Cfg<PAIR>;
  constexpr int AST = C::A_STAGES, BST = C::B_STAGES, B_TILE = C::B_TILE, B_STAGE = C::B_STAGE, NC = PAIR ? 2 : 1;
  constexpr bool r_tma = RTMA != 0;
  launch_dependents();
  extern __shared__ uint8_t og_tcf_smem_raw[];
  uint8_t* smem = tc::align_smem_1024(og_tcf_smem_raw);
  uint8_t* sm_a = smem;                                                  // [AST][32 KB] raw fp32 A tiles
  uint8_t* sm_b = smem + AST * A_BYTES;                                  // [BST][B_hi | B_lo]
  uint8_t* s_out = sm_b + BST * B_STAGE;                                  // [2 warpgroups][2 buffers][16 KB]
  Barriers* bars = reinterpret_cast<Barriers*>(s_out + OUT_BYTES);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int g_first = blockIdx.x / NC, g_stride = gridDim.x / NC;

  if (threadIdx.x == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(&bars->a_land[i], 1); mbar_init(&bars->a_free[i], 4);      // 4 converter warps have read the smem tile
      mbar_init(&bars->a_full[i], 4 * NC); mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < BST; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 8 * NC); mbar_init(&bars->r_full[i], 1); }
    fence_barrier_init();
    prefetch_tensormap(&map_a); prefetch_tensormap(&map_a2);
    prefetch_tensormap(&map_bhi); prefetch_tensormap(&map_blo);
  }
  if (PAIR) cluster_sync_all();
  if (warp == 13) { if (PAIR) tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base); else tmem_alloc<TMEM_COLS>(&bars->tmem_base); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  grid_dependency_wait();                                    // the amax slots below are written by the previous kernels

  // operand scales: every thread derives the same values from the same device scalars
  float amax_a = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) if (a.amax_in[i]) amax_a = fmaxf(amax_a, __ldcg(a.amax_in[i]));
  const float s_a = f16_scale_for(amax_a);
  const float s_w = __ldg(a.w_meta);
  const float alpha_eff = a.alpha / (s_a * s_w);             // exact: both scales are powers of two
  float s_out_scale = 1.f;
  if (OUTK == 2 || OUTK == 3) {
    s_out_scale = f16_scale_for(fabsf(a.alpha) * fmaf(amax_a, __ldg(a.w_meta + 1), __ldg(a.w_meta + 2)));
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.scale_out) *a.scale_out = s_out_scale;
  }
  if (OUTK == 4 && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int kk = 0; kk < a.nkinds; ++kk) {
      const float* wm = a.w_meta + 4 * kk;
      const float so = f16_scale_for(fabsf(a.alpha) * fmaf(amax_a, __ldg(wm + 1), __ldg(wm + 2)));
      if (a.kind0 + kk == 1 && a.scale_out) *a.scale_out = so;
      if (a.kind0 + kk == 2 && a.scale_out_v) *a.scale_out_v = so;
    }
  }

  auto arrive_leader = [&](uint64_t* bar) {
    __syncwarp();
    if (lane == 0) { if (!PAIR || crank == 0) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); }
  };
  auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
  auto tile_coords = [&](int t, int& m0, int& n0, int& bz) {
    n0 = (t % sc.ntn) * BN;
    m0 = (((t / sc.ntn) % sc.ntmg) * NC + (int)crank) * BM;
    bz = t / (sc.ntn * sc.ntmg);
  };

  if (warp >= 12) {
  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer of the fp32 A tiles
    if (elect_one()) {
      int it = 0;
      for (int t = g_first; t < sc.ngroups; t += g_stride) {
        int m0, n0, bz; tile_coords(t, m0, n0, bz);
        const int arow = bz * a.rows + m0;
        for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
          const int s = it % AST, ph = (it / AST) & 1;
          mbar_wait(&bars->a_free[s], ph ^ 1);
          OG_TRACE_EVT(0, it);
          uint8_t* dst = sm_a + s * A_BYTES;
          const int k = kb * BK;
          mbar_arrive_expect_tx(&bars->a_land[s], A_BYTES);    // columns beyond K arrive as zeros (TMA out-of-bounds fill)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int kj = k + 32 * j;
            if (kj < a.k1 || !a.A2) tma_load_2d(dst + j * A_SUB, &map_a, &bars->a_land[s], kj, arow);      // kj >= K: all zeros
            else                    tma_load_2d(dst + j * A_SUB, &map_a2, &bars->a_land[s], kj - a.k1, arow);
          }
        }
      }
    }
  } else if (warp == 14) {
    // ------------------------------------------------------------------ TMA producer of the fp16 B tiles (own, deeper ring)
    if (elect_one()) {
      int it = 0;
      for (int t = g_first; t < sc.ngroups; t += g_stride) {
        int m0, n0, bz; tile_coords(t, m0, n0, bz);
        const int brow = n0 + bz * a.b_rows_per_batch + (int)crank * C::BROWS;
        for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
          const int s = it % BST, ph = (it / BST) & 1;
          mbar_wait(&bars->b_free[s], ph ^ 1);
          uint8_t* dst = sm_b + s * B_STAGE;
          const int k = kb * BK;
          if (crank == 0) mbar_arrive_expect_tx(&bars->b_full[s], NC * B_STAGE);
          if (PAIR) {
            tma_load_2d_pair(dst, &map_bhi, &bars->b_full[s], k, brow);
            tma_load_2d_pair(dst + B_TILE, &map_blo, &bars->b_full[s], k, brow);
          } else {
            tma_load_2d(dst, &map_bhi, &bars->b_full[s], k, brow);
            tma_load_2d(dst + B_TILE, &map_blo, &bars->b_full[s], k, brow);
          }
        }
      }
    }
  } else if (warp == 13 && crank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only when paired)
    const uint32_t idesc = make_idesc_f16(BM * NC, BN);
    int it = 0, g = 0;                                         // k-block and chunk counters
    for (int t = g_first; t < sc.ngroups; t += g_stride) {
      for (int c = 0; c < sc.nchunks; ++c, ++g) {
        const int buf = g & 1, gph = (g >> 1) & 1;
        mbar_wait(&bars->acc_empty[buf], gph ^ 1);
        tc_fence_after();
        const int kb_end = min((c + 1) * CHUNK_KB, sc.nkb);
        for (int kb = c * CHUNK_KB; kb < kb_end; ++kb, ++it) {
          const int s = it % AST, ph = (it / AST) & 1;           // A: TMEM ring slot
          const int sb = it % BST, phb = (it / BST) & 1;         // B: smem ring slot
          mbar_wait(&bars->b_full[sb], phb);
          OG_TRACE_EVT(3, it);
          mbar_wait(&bars->a_full[s], ph);
          tc_fence_after();
          OG_TRACE_EVT(4, it);
          if (elect_one()) {
            const uint32_t bhi = smem_u32(sm_b + sb * B_STAGE), blo = bhi + B_TILE;
            const uint32_t d = tmem + buf * 128;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t dbhi = make_sdesc_sw128(bhi + kk * 32), dblo = make_sdesc_sw128(blo + kk * 32);
              const uint32_t ahi = tmem + COL_A + s * 64 + kk * 8, alo = ahi + 32;
              const uint32_t acc0 = (kb > c * CHUNK_KB || kk) ? 1u : 0u;
              if (PAIR) {
                umma_f16_ts_pair(d, alo, dbhi, idesc, acc0);
                umma_f16_ts_pair(d, ahi, dblo, idesc, 1u);
                umma_f16_ts_pair(d, ahi, dbhi, idesc, 1u);
              } else {
                umma_f16_ts(d, alo, dbhi, idesc, acc0);
                umma_f16_ts(d, ahi, dblo, idesc, 1u);
                umma_f16_ts(d, ahi, dbhi, idesc, 1u);
              }
            }
            commit(&bars->b_free[sb]);
            commit(&bars->a_empty[s]);
            if (kb == kb_end - 1) commit(&bars->acc_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ A converters: smem fp32 -> scale, split, pack -> TMEM
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int t = g_first; t < sc.ngroups; t += g_stride) {
      for (int kb = 0; kb < sc.nkb; ++kb, ++it) {
        const int s = it % AST, ph = (it / AST) & 1;
        mbar_wait(&bars->a_land[s], ph);
        if (warp == 8 && lane == 0) OG_TRACE_EVT(1, it);
        uint32_t hi[32], lo[32];
#if OG_GEMM_CONV2
        // the whole row (64 floats) goes to registers first and the smem slot is handed back BEFORE the split: a slot lives
        // TMA latency (~2500 cycles) + the time until this arrive, and the A ring is what paces the main loop (3 slots of 32 KB;
        // event trace profiles/r02_trace_gemm_f16_fc2_v2_two_rings.txt: one K block per ~1150 cycles against 768 of MMA work).
        // The loads are ordered before the release-arrive by the memory model (__syncwarp orders the lanes' accesses).
        ulonglong2 v[16];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint8_t* arow = sm_a + s * A_BYTES + j * A_SUB + trow * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) v[j * 8 + c] = *reinterpret_cast<const ulonglong2*>(arow + ((c ^ (trow & 7)) * 16));   // undo the 128B swizzle
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->a_free[s]);
        {
          const unsigned long long s2 = tcf::pk2(s_a, s_a);
#pragma unroll
          for (int w = 0; w < 16; ++w) {               // packed fp32x2 scale / residual (bit-identical to the scalar form)
            tcf::split_f16x2_packed(tcf::mul2(v[w].x, s2), hi[2 * w], lo[2 * w]);
            tcf::split_f16x2_packed(tcf::mul2(v[w].y, s2), hi[2 * w + 1], lo[2 * w + 1]);
          }
        }
        if (a.swap_halves) {
#pragma unroll
          for (int w = 0; w < 32; ++w) { hi[w] = __byte_perm(hi[w], 0, 0x1032); lo[w] = __byte_perm(lo[w], 0, 0x1032); }
        }
#else
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint8_t* arow = sm_a + s * A_BYTES + j * A_SUB + trow * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (trow & 7)) * 16));   // undo the 128B swizzle
            split_f16x2(v.x * s_a, v.y * s_a, hi[j * 16 + 2 * c], lo[j * 16 + 2 * c]);
            split_f16x2(v.z * s_a, v.w * s_a, hi[j * 16 + 2 * c + 1], lo[j * 16 + 2 * c + 1]);
          }
        }
        if (a.swap_halves) {
#pragma unroll
          for (int w = 0; w < 32; ++w) { hi[w] = __byte_perm(hi[w], 0, 0x1032); lo[w] = __byte_perm(lo[w], 0, 0x1032); }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->a_free[s]);          // this warp is done with the smem A tile: the producer may refill it
#endif
        mbar_wait(&bars->a_empty[s], ph ^ 1);
        tc_fence_after();
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
    const int half = warp >> 2;
    constexpr int HN = BN / 2;
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int wg_tid = threadIdx.x & 127;
    int g = 0, ntile = 0;
    float tmax = 0.f;                                          // running max |y| of this thread (OUTK 1, amax_out)
    for (int t = g_first; t < sc.ngroups; t += g_stride, ++ntile) {
      int m0, n0, bz; tile_coords(t, m0, n0, bz);
      float racc[HN];
#pragma unroll
      for (int j = 0; j < HN; ++j) racc[j] = 0.f;
      const int pb = ntile & 1;
      if (wg_tid < 64) {
        const int c = half * HN + wg_tid, col = n0 + c;
        bars->bias[pb][c] = (a.bias && col < a.nout) ? __ldg(a.bias + col) : 0.f;
      }
      if (r_tma && wg_tid == 0) {
        tma_store_wait_read<0>();
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
      const int grow = m0 + trow;
      const bool row_ok = grow < a.rows;
      // output kind of this tile, its column origin inside that output tensor, and the kind's operand scales
      int okind = OUTK, kidx = 0, ocols = a.nout;
      float alpha_t = alpha_eff, sos = s_out_scale;
      if (OUTK == 4) {
        kidx = n0 / a.kind_cols; okind = a.kind0 + kidx + 1; ocols = a.kind_cols;
        const float* wm = a.w_meta + 4 * kidx;
        alpha_t = a.alpha / (s_a * __ldg(wm));
        if (okind != 1) sos = f16_scale_for(fabsf(a.alpha) * fmaf(amax_a, __ldg(wm + 1), __ldg(wm + 2)));
      }
      if (wg_tid == 0 && !r_tma && OUTK != 3) tma_store_wait_read<0>();   // the previous tile's stores have read the staging buffers
      asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");       // ... and the staged bias is visible
      const int cl0 = half * HN, cb0 = n0 + cl0;               // first column of this warpgroup's half (tile / stacked B rows)
      const int ob0 = cb0 - kidx * ocols;                      // ... inside its output tensor
      if (cb0 < a.nout) {                                      // uniform across the warpgroup
        float y[HN];
#pragma unroll
        for (int j = 0; j < HN; j += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&bars->bias[pb][cl0 + j]);
          y[j] = fmaf(racc[j], alpha_t, bv.x);         y[j + 1] = fmaf(racc[j + 1], alpha_t, bv.y);
          y[j + 2] = fmaf(racc[j + 2], alpha_t, bv.z); y[j + 3] = fmaf(racc[j + 3], alpha_t, bv.w);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < HN; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (okind == 1) {
          if (r_tma) mbar_wait(&bars->r_full[half], ntile & 1);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {                     // two 32-column fp32 chunks, each staged in its own buffer
            uint8_t* buf = s_out + (half * 2 + cc) * OUT_TILE + trow * 128;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              float4 o = make_float4(y[cc * 32 + 4 * c4], y[cc * 32 + 4 * c4 + 1], y[cc * 32 + 4 * c4 + 2], y[cc * 32 + 4 * c4 + 3]);
              float4* cell = reinterpret_cast<float4*>(buf + ((c4 ^ (trow & 7)) * 16));
              if (r_tma) { const float4 r = *cell; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
              *cell = o;
              if (a.amax_out && row_ok)                        // columns beyond nout hold alpha * 0 + 0 = 0: harmless for a max
                tmax = fmaxf(tmax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
            }
          }
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
          if (wg_tid == 0) {
            tma_store_3d(&map_y, s_out + (half * 2 + 0) * OUT_TILE, ob0, m0, bz);
            if (ob0 + 32 < ocols) tma_store_3d(&map_y, s_out + (half * 2 + 1) * OUT_TILE, ob0 + 32, m0, bz);
            tma_store_commit();
          }
        } else if (okind == 2) {
          uint8_t* bh = s_out + (half * 2 + 0) * OUT_TILE + trow * 128;      // hi row: 64 halves = 128 bytes
          uint8_t* bl = s_out + (half * 2 + 1) * OUT_TILE + trow * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {                        // 16-byte chunk c = columns 8c .. 8c+7
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_f16x2(y[8 * c + 2 * e] * sos, y[8 * c + 2 * e + 1] * sos, h[e], l[e]);
            *reinterpret_cast<uint4*>(bh + ((c ^ (trow & 7)) * 16)) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(bl + ((c ^ (trow & 7)) * 16)) = make_uint4(l[0], l[1], l[2], l[3]);
          }
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
          if (wg_tid == 0) {
            tma_store_3d(&map_yh, s_out + (half * 2 + 0) * OUT_TILE, ob0, m0, bz);
            tma_store_3d(&map_yl, s_out + (half * 2 + 1) * OUT_TILE, ob0, m0, bz);
            tma_store_commit();
          }
        } else {                                               // kind 3: for a fixed column the 32 lanes write 32 consecutive rows
          if (row_ok) {
            const int64_t ytoff = (int64_t)bz * a.strideYt + grow;
#pragma unroll
            for (int j = 0; j < HN; j += 2) {
              uint32_t h, l;
              split_f16x2(y[j] * sos, y[j + 1] * sos, h, l);
              if (ob0 + j < ocols) {
                const int64_t o = ytoff + (int64_t)(ob0 + j) * a.ldyt;
                a.Yth[o] = __ushort_as_half((unsigned short)(h & 0xffffu)); a.Ytl[o] = __ushort_as_half((unsigned short)(l & 0xffffu));
              }
              if (ob0 + j + 1 < ocols) {
                const int64_t o = ytoff + (int64_t)(ob0 + j + 1) * a.ldyt;
                a.Yth[o] = __ushort_as_half((unsigned short)(h >> 16)); a.Ytl[o] = __ushort_as_half((unsigned short)(l >> 16));
              }
            }
          }
        }
      }
      if (warp == 0 && lane == 0) OG_TRACE_EVT(9, ntile);
    }
    if ((OUTK == 1 || OUTK == 4) && a.amax_out) {
      tmax = warp_max(tmax);
      if (lane == 0 && tmax > 0.f) atomic_amax(a.amax_out, tmax);
    }
    if (wg_tid == 0 && OUTK != 3) tma_store_wait_all<0>();      // smem must outlive the last store's reads
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 13) { tc_fence_after(); if (PAIR) tmem_dealloc_pair<tcf::TMEM_COLS>(tmem); else tmem_dealloc<tcf::TMEM_COLS>(tmem); }
}

inline bool linear_f16_eligible(const F16LinearArgs& a, const __half* Bh, const __half* Bl, int64_t ldb) {
  const int K = a.k1 + a.k2;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!(K >= 64 && a.k1 % 4 == 0 && a.k2 % 4 == 0 && a.lda % 4 == 0 && al16(a.A) && ldb % 8 == 0 && al16(Bh) && al16(Bl))) return false;
  if (a.batch > 1 && a.strideA != (int64_t)a.rows * a.lda) return false;            // dense batches (one 2-D tensor map)
  if (a.A2 && (a.k1 % 64 != 0 || a.lda2 % 4 != 0 || !al16(a.A2) || (a.batch > 1 && a.strideA2 != (int64_t)a.rows * a.lda2))) return false;
  if (!a.w_meta || !(a.amax_in[0] || a.amax_in[1] || a.amax_in[2])) return false;
  const int kinds = (a.Y ? 1 : 0) + (a.Yh ? 1 : 0) + (a.Yth ? 1 : 0);
  if (a.nkinds > 1) {                                       // stacked projections: kinds kind0 .. kind0 + nkinds - 1, one output each
    if (a.kind0 < 0 || a.kind0 + a.nkinds > 3 || kinds != a.nkinds || a.kind_cols % tcf::BN != 0 || a.nout != a.nkinds * a.kind_cols) return false;
    if ((a.kind0 == 0) != (a.Y != nullptr) || !a.Yh || (a.kind0 + a.nkinds == 3) != (a.Yth != nullptr) || a.R || a.relu) return false;
    if (a.Yth && !a.scale_out_v) return false;
  } else if (kinds != 1) return false;
  if (a.Y && !(a.ldy % 4 == 0 && a.strideY % 4 == 0 && al16(a.Y))) return false;
  if (a.Y && a.R && !(a.ldr % 4 == 0 && a.strideR % 4 == 0 && al16(a.R) && a.nout % 32 == 0)) return false;
  if (a.Yh && !(a.Yl && a.ldy % 8 == 0 && a.strideY % 8 == 0 && al16(a.Yh) && al16(a.Yl) && !a.R && a.scale_out)) return false;
  if (a.Yth && !(a.Ytl && !a.R && (a.scale_out || a.scale_out_v))) return false;
  return true;
}

template <int PAIR>
inline int linear_f16_launch_t(const F16LinearArgs& a, const __half* Bh, const __half* Bl, int64_t ldb, int64_t b_total_rows,
                               cudaStream_t stream) {
  using namespace tcf;
  using C = Cfg<PAIR>;
  constexpr int NC = PAIR ? 2 : 1;
  const int K = a.k1 + a.k2;
  CUtensorMap ma, ma2, mh, ml;
  int rc;
  const uint64_t arows = (uint64_t)a.batch * a.rows;
  if ((rc = tc::make_tmap_2d(&ma, a.A, arows, (uint64_t)a.k1, (uint64_t)a.lda, BM)) != OG_OK) return rc;
  if (a.A2) { if ((rc = tc::make_tmap_2d(&ma2, a.A2, arows, (uint64_t)a.k2, (uint64_t)a.lda2, BM)) != OG_OK) return rc; }
  else ma2 = ma;
  if ((rc = tc::make_tmap_2d_f16(&mh, Bh, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, C::BROWS)) != OG_OK) return rc;
  if ((rc = tc::make_tmap_2d_f16(&ml, Bl, (uint64_t)b_total_rows, (uint64_t)K, (uint64_t)ldb, C::BROWS)) != OG_OK) return rc;
  CUtensorMap my = ma, myh = ma, myl = ma, mr = ma;
  const int ocols = a.nkinds > 1 ? a.kind_cols : a.nout;     // columns of ONE output tensor
  if (a.Y && (rc = tc::make_tmap_3d(&my, a.Y, a.batch, a.rows, ocols, a.ldy, a.strideY, BM)) != OG_OK) return rc;
  if (a.Yh && (rc = tc::make_tmap_3d_f16(&myh, a.Yh, a.batch, a.rows, ocols, a.ldy, a.strideY, BM)) != OG_OK) return rc;
  if (a.Yh && (rc = tc::make_tmap_3d_f16(&myl, a.Yl, a.batch, a.rows, ocols, a.ldy, a.strideY, BM)) != OG_OK) return rc;
  const int r_tma = a.Y && a.R;
  if (r_tma && (rc = tc::make_tmap_3d(&mr, a.R, a.batch, a.rows, a.nout, a.ldr, a.strideR, BM)) != OG_OK) return rc;
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                        F16LinearArgs, Sched);
  static const Kern kerns[5] = {linear_f16_kernel<PAIR, 0, 1>, linear_f16_kernel<PAIR, 1, 1>, linear_f16_kernel<PAIR, 0, 2>,
                                linear_f16_kernel<PAIR, 0, 3>, linear_f16_kernel<PAIR, 0, 4>};
  static DeviceFlags attr_set;
  if (attr_set.once()) {
    for (Kern k : kerns) OG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  const Kern kern = a.nkinds > 1 ? kerns[4] : a.Y ? (r_tma ? kerns[1] : kerns[0]) : (a.Yh ? kerns[2] : kerns[3]);
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
  OG_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, ma2, mh, ml, my, myh, myl, mr, a, sc));
  launch_counter()++;
  return OG_OK;
}

inline int linear_f16_launch(const F16LinearArgs& a, const __half* Bh, const __half* Bl, int64_t ldb, int64_t b_total_rows,
                             cudaStream_t stream) {
  const bool pair = linear_tc2_pair_mode() != 0 && cdiv(a.rows, tcf::BM) >= 2;
  return pair ? linear_f16_launch_t<1>(a, Bh, Bl, ldb, b_total_rows, stream)
              : linear_f16_launch_t<0>(a, Bh, Bl, ldb, b_total_rows, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// pack-time / set-up kernels
// amax of a flat fp32 buffer into a device slot (the GNN's input activations come out of the CUDA-core encoder layers)
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ slot) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(x + i)));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) tc::atomic_amax(slot, m);
}
// weight tensor [rows, cols] (+ bias [rows]) -> meta = {scale, max_n ||W_n||_1, max |b|}; one warp per row, atomics on
// non-negative floats (order independent); meta must be zeroed first.  meta[0] temporarily holds amax(W).
__global__ void __launch_bounds__(256) weight_meta_kernel(const float* __restrict__ w, const float* __restrict__ bias, int rows, int cols,
                                                          float* __restrict__ meta) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float mx = 0.f, l1 = 0.f;
  for (int c = lane; c < cols; c += 32) { const float v = fabsf(__ldg(w + (int64_t)row * cols + c)); mx = fmaxf(mx, v); l1 += v; }
  mx = warp_max(mx); l1 = warp_sum(l1);
  if (lane == 0) {
    tc::atomic_amax(meta, mx); tc::atomic_amax(meta + 1, l1 * 1.0001f);           // 1.0001: the sum itself is rounded
    if (bias) tc::atomic_amax(meta + 2, fabsf(__ldg(bias + row)));
  }
}
// x -> hi / lo halves with the tensor's scale (meta[0] holds amax(W) on entry; the LAST block to finish replaces it by the scale)
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                                                        int64_t n, const float* __restrict__ amax) {
  const float s = tc::f16_scale_for(__ldg(amax));
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float x = src[i] * s;
    const __half h = __float2half_rn(x);
    hi[i] = h; lo[i] = __float2half_rn(x - __half2float(h));
  }
}
__global__ void finish_meta_kernel(float* meta) { meta[0] = tc::f16_scale_for(meta[0]); }

}  // namespace og
