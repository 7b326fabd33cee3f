// extern "C" entry points of libopenglue_b200.so (see include/openglue_b200.h) and the
// host-side schedule of the whole matching-core forward pass.
#include "common.cuh"
#include "linear_simt.cuh"
#include "linear_tc.cuh"
#include "linear_tc2.cuh"
#include "attention_simt.cuh"
#include "attention_tc.cuh"
#include "linear_f16.cuh"
#include "attention_f16.cuh"
#include "attention_f16t.cuh"
#include "attention_f16p.cuh"
#include "sinkhorn.cuh"
#include "sinkhorn_bwd.cuh"
#include "match.cuh"
#include "gt_matches.cuh"
#include "criterion.cuh"
#include "collate.cuh"
#include "train_ops.cuh"
#include "superpoint.cuh"
#include <math.h>
#include <string.h>
#include <vector>

namespace og {

// keypoint normalisation + concat with side info:  in0[r] = [2*x/(W-1) - 1, 2*y/(H-1) - 1, side...]
// (reference superglue.py:74-78 and positional_encoding.py:16-18)
__global__ void __launch_bounds__(256) kenc_input_kernel(const float* __restrict__ kpts, const float* __restrict__ side,
                                                          int rows, int S, float wm1, float hm1, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float* o = out + (int64_t)r * (2 + S);
  o[0] = __fdiv_rn(2.f * __ldg(kpts + 2 * (int64_t)r), wm1) - 1.f;
  o[1] = __fdiv_rn(2.f * __ldg(kpts + 2 * (int64_t)r + 1), hm1) - 1.f;
  for (int s = 0; s < S; ++s) o[2 + s] = __ldg(side + (int64_t)r * S + s);
}

struct Layout {           // float offsets of the packed weights
  std::vector<int64_t> kenc_w, kenc_b;
  std::vector<int64_t> qkv_w, qkv_b, fc1_w, fc1_b, fc2_w, fc2_b;
  int64_t proj_w, proj_b, proj_rmix, dustbin, total;
  std::vector<int> kenc_sizes;
};

static int check_config(const og_config* c) {
  OG_CHECK_ARG(c != nullptr, "config is null");
  OG_CHECK_ARG(c->descriptor_dim > 0 && c->num_heads > 0 && c->descriptor_dim % c->num_heads == 0,
               "descriptor_dim %d must be a positive multiple of num_heads %d", c->descriptor_dim, c->num_heads);
  OG_CHECK_ARG(c->num_layers >= 0 && c->num_layers % 2 == 0, "num_layers must be 2 * num_stages");
  OG_CHECK_ARG(c->side_info_size >= 0, "side_info_size < 0");
  OG_CHECK_ARG(c->num_hidden >= 0 && c->num_hidden <= OG_MAX_HIDDEN, "num_hidden out of range");
  OG_CHECK_ARG(c->sinkhorn_iters >= 0 && c->sinkhorn_reg > 0.f, "bad sinkhorn parameters");
  OG_CHECK_ARG(c->descriptor_dim % 4 == 0, "descriptor_dim must be a multiple of 4");
  return OG_OK;
}

static Layout make_layout(const og_config* c) {
  Layout L;
  const int64_t d = c->descriptor_dim;
  L.kenc_sizes.push_back(2 + c->side_info_size);
  for (int i = 0; i < c->num_hidden; ++i) L.kenc_sizes.push_back(c->hidden[i]);
  L.kenc_sizes.push_back((int)d);
  int64_t off = 0;
  auto take = [&](int64_t nfl) { int64_t o = off; off += (nfl + 3) / 4 * 4; return o; };   // keep 16-byte alignment
  for (size_t i = 1; i < L.kenc_sizes.size(); ++i) {
    L.kenc_w.push_back(take((int64_t)L.kenc_sizes[i] * L.kenc_sizes[i - 1]));
    L.kenc_b.push_back(take(L.kenc_sizes[i]));
  }
  for (int l = 0; l < c->num_layers; ++l) {
    L.qkv_w.push_back(take(3 * d * d)); L.qkv_b.push_back(take(3 * d));
    L.fc1_w.push_back(take(4 * d * d)); L.fc1_b.push_back(take(2 * d));
    L.fc2_w.push_back(take(2 * d * d)); L.fc2_b.push_back(take(d));
  }
  L.proj_w = take(d * d); L.proj_b = take(d); L.proj_rmix = take(d); L.dustbin = take(1);
  L.total = off;
  return L;
}

struct Workspace {
  float *in0, *h0, *h1, *x, *qkv, *o, *hid, *g, *ghi, *glo, *sbuf;
  float *q, *khi, *klo, *vthi, *vtlo;     // tensor-core attention operands (OG_PREC_TF32X3)
  int64_t ldn, ldm;                       // padded row lengths of the channel-major V^T buffers
  void *sink, *match;
  float* slots;                           // amax / scale scalars of the fp16 path (zeroed at the start of a forward pass)
  int nslots;
  int64_t lds, sink_bytes, match_bytes, total;
};

static int plan_workspace(const og_config* c, int B, int n, int m, void* base, Workspace* w) {
  const int64_t d = c->descriptor_dim, R = (int64_t)B * (n + m);
  int maxh = (int)d;
  for (int i = 0; i < c->num_hidden; ++i) maxh = std::max(maxh, c->hidden[i]);
  char* p = static_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += align_up(bytes, 256); return p ? (void*)(p + o) : nullptr; };
  w->in0 = (float*)take(R * (2 + c->side_info_size) * 4);
  w->h0 = (float*)take(R * maxh * 4);
  w->h1 = (float*)take(R * maxh * 4);
  w->x = (float*)take(R * d * 4);
  w->qkv = (float*)take(R * 3 * d * 4);
  w->o = (float*)take(R * d * 4);
  w->hid = (float*)take(R * 2 * d * 4);
  w->g = (float*)take(R * d * 4);
  w->ghi = (float*)take((int64_t)B * m * d * 4);          // tf32 split of image-1 descriptors (score GEMM B operand)
  w->glo = (float*)take((int64_t)B * m * d * 4);
  w->ldn = align_up(n, 8); w->ldm = align_up(m, 8);          // 16-byte rows for the fp16 form as well
  w->nslots = 16 * c->num_layers + 16;
  w->slots = (float*)take((int64_t)w->nslots * 4);
  if (c->precision == OG_PREC_TF32X3 || c->precision == OG_PREC_FP16X3) {
    w->q = (float*)take(R * d * 4);
    w->khi = (float*)take(R * d * 4);
    w->klo = (float*)take(R * d * 4);
    w->vthi = (float*)take((int64_t)B * d * (w->ldn + w->ldm) * 4);
    w->vtlo = (float*)take((int64_t)B * d * (w->ldn + w->ldm) * 4);
  } else {
    w->q = w->khi = w->klo = w->vthi = w->vtlo = nullptr;
  }
  w->lds = align_up(m, 4);
  w->sbuf = (float*)take((int64_t)B * n * w->lds * 4);
  w->sink_bytes = sinkhorn_workspace_bytes(B, n, m);
  if (w->sink_bytes < 0) return OG_EUNSUPPORTED;
  w->sink = take(w->sink_bytes);
  w->match_bytes = match_workspace_bytes(B, n, m);
  w->match = take(w->match_bytes);
  w->total = off;
  return OG_OK;
}

static og_linear_args lin(const float* A, int64_t lda, int k, const float* W, const float* bias, int rows, int nout,
                          float* Y, int64_t ldy) {
  og_linear_args a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = lda; a.k1 = k; a.W = W; a.ldw = k; a.bias = bias; a.rows = rows; a.nout = nout; a.batch = 1;
  a.alpha = 1.f; a.Y = Y; a.ldy = ldy;
  return a;
}

static TcLinearArgs to_tc_args(const og_linear_args& a) {
  TcLinearArgs t;
  memset(&t, 0, sizeof(t));
  t.A = a.A; t.lda = a.lda; t.strideA = a.strideA; t.A2 = a.A2; t.lda2 = a.lda2; t.strideA2 = a.strideA2;
  t.k1 = a.k1; t.k2 = a.k2;
  t.b_rows_per_batch = a.strideW ? (int)(a.strideW / a.ldw) : 0;
  t.bias = a.bias; t.rows = a.rows; t.nout = a.nout; t.batch = a.batch; t.alpha = a.alpha; t.relu = a.relu;
  t.R = a.R; t.ldr = a.ldr; t.strideR = a.strideR; t.rscale = a.rscale;
  t.Y = a.Y; t.ldy = a.ldy; t.strideY = a.strideY; t.Yt = a.Yt; t.ldyt = a.ldyt; t.strideYt = a.strideYt;
  return t;
}

// Split-output request for the tensor-core path (the fp32 CUDA-core path ignores it).
struct SplitOut { float *Yhi = nullptr, *Ylo = nullptr, *Ythi = nullptr, *Ytlo = nullptr; };

static int linear_tc_run(const og_linear_args& a, const float* Whi, const float* Wlo, const SplitOut& so, int mode,
                         cudaStream_t s) {
  TcLinearArgs t = to_tc_args(a);
  t.Yhi = so.Yhi; t.Ylo = so.Ylo; t.Ythi = so.Ythi; t.Ytlo = so.Ytlo;
  if (a.strideW && a.strideW % a.ldw != 0) return fail(OG_EUNSUPPORTED, "linear_tc: strideW must be a multiple of ldw");
  if (!linear_tc_eligible(t, Whi, Wlo, a.ldw))
    return fail(OG_EUNSUPPORTED, "linear_tc: needs K >= 32, K %% 4 == 0 and 16-byte aligned rows");
  // rows the B tensor map may touch: the LAST batch item only owns nout rows (a map declared over b_rows_per_batch * batch rows
  // would let a 128-row TMA box read past the end of a head-sliced or exactly-sized operand; rows beyond the map are zero-filled)
  const int64_t brows = a.strideW ? (int64_t)t.b_rows_per_batch * (a.batch - 1) + a.nout : a.nout;
  if (mode == 2) {                       // production kernel: persistent, chunked accumulation
    if (!linear_tc2_eligible(t, Whi, Wlo, a.ldw)) return fail(OG_EUNSUPPORTED, "linear_tc2: needs dense batches and k1 %% 32 == 0 for concat");
    return linear_tc2_launch(t, Whi, Wlo, a.ldw, brows, s);
  }
  return mode == tcl::MODE_SS ? linear_tc_launch_mode<tcl::MODE_SS>(t, Whi, Wlo, a.ldw, brows, s)
                              : linear_tc_launch_mode<tcl::MODE_TS>(t, Whi, Wlo, a.ldw, brows, s);
}

// Kernel selection for one linear layer: tcgen05 3xTF32 when asked for and the shape is tileable,
// otherwise the fp32 CUDA-core kernel (tiny K such as the 3-channel keypoint-encoder input).
static int linear_dispatch(const og_linear_args& a, int precision, cudaStream_t s, const float* Whi = nullptr,
                           const float* Wlo = nullptr, const SplitOut& so = SplitOut()) {
  if (precision == OG_PREC_TF32X3 && Whi && Wlo) {
    TcLinearArgs t = to_tc_args(a);
    if (linear_tc2_eligible(t, Whi, Wlo, a.ldw)) return linear_tc_run(a, Whi, Wlo, so, 2, s);
    if (linear_tc_eligible(t, Whi, Wlo, a.ldw)) return linear_tc_run(a, Whi, Wlo, so, tcl::MODE_TS, s);
  }
  if (so.Yhi || so.Ythi) return fail(OG_EUNSUPPORTED, "split outputs need the tensor-core path");
  return linear_simt_launch(a, s);
}

// exact fp32 attention (CUDA cores); the tensor-core forms take pre-split operands and have their own entry points
static int attention_fp32(const AttnArgs& a, int head_dim, cudaStream_t s) { return attention_simt_launch(a, head_dim, s); }

}  // namespace og

using namespace og;

extern "C" {

int og_version(void) { return OG_VERSION; }
const char* og_last_error(void) { return err_buf(); }

int og_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  const DeviceInfo& d = device_info();
  if (!d.ok) return fail(OG_ECUDA, "no CUDA device");
  if (sm_count) *sm_count = d.sm_count;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  return OG_OK;
}

int64_t og_packed_weight_floats(const og_config* cfg) {
  if (check_config(cfg) != OG_OK) return -1;
  return make_layout(cfg).total;
}

int64_t og_packed_offset(const og_config* cfg, int tensor_id, int index) {
  if (check_config(cfg) != OG_OK) return -1;
  Layout L = make_layout(cfg);
  auto at = [&](const std::vector<int64_t>& v) -> int64_t {
    return (index >= 0 && index < (int)v.size()) ? v[index] : (int64_t)fail(OG_EINVAL, "index %d out of range", index);
  };
  switch (tensor_id) {
    case OG_T_KENC_W: return at(L.kenc_w);
    case OG_T_KENC_B: return at(L.kenc_b);
    case OG_T_QKV_W: return at(L.qkv_w);
    case OG_T_QKV_B: return at(L.qkv_b);
    case OG_T_FC1_W: return at(L.fc1_w);
    case OG_T_FC1_B: return at(L.fc1_b);
    case OG_T_FC2_W: return at(L.fc2_w);
    case OG_T_FC2_B: return at(L.fc2_b);
    case OG_T_PROJ_W: return L.proj_w;
    case OG_T_PROJ_B: return L.proj_b;
    case OG_T_PROJ_RMIX: return L.proj_rmix;
    case OG_T_DUSTBIN: return L.dustbin;
    default: return fail(OG_EINVAL, "unknown tensor id %d", tensor_id);
  }
}

int64_t og_workspace_bytes(const og_config* cfg, int batch, int n, int m) {
  if (check_config(cfg) != OG_OK) return -1;
  if (batch <= 0 || n <= 0 || m <= 0) return fail(OG_EINVAL, "batch, n, m must be positive");
  Workspace w;
  if (plan_workspace(cfg, batch, n, m, nullptr, &w) != OG_OK) return -1;
  return w.total;
}

int og_last_forward_launches(void) { return launch_counter(); }

static int& fuse_qkv_mode() {
  static int v = [] { const char* e = getenv("OG_FUSE_QKV"); return e ? (atoi(e) != 0) : 1; }();
  return v;
}
int og_set_fusion(int fuse_projections) {
  const int prev = fuse_qkv_mode();
  if (fuse_projections >= 0) fuse_qkv_mode() = fuse_projections ? 1 : 0;
  return prev;
}

int og_set_tuning(int gemm_pair, int attention_pair) {
  if (gemm_pair >= 0) linear_tc2_pair_mode() = gemm_pair ? 1 : 0;
  if (attention_pair >= 0) attention_tc_pair_mode() = attention_pair ? 1 : 0;
  return OG_OK;
}

int og_linear_fwd(const og_linear_args* a, int precision, void* stream) {
  OG_CHECK_ARG(a && a->A && a->W && (a->Y || a->Yt), "linear: null pointer");
  OG_CHECK_ARG(precision == OG_PREC_FP32, "linear: the tensor-core form takes pre-split weights (og_linear_tc_fwd)");
  OG_CHECK_ARG(a->rows > 0 && a->nout > 0 && a->batch > 0 && a->k1 > 0 && a->k2 >= 0, "linear: bad sizes");
  OG_CHECK_ARG(a->k2 == 0 || a->A2, "linear: k2 > 0 needs A2");
  return linear_dispatch(*a, precision, (cudaStream_t)stream);
}

int og_linear_tc_fwd(const og_linear_args* a, const float* Whi, const float* Wlo, float* Yhi, float* Ylo, float* Ythi,
                     float* Ytlo, int mode, void* stream) {
  OG_CHECK_ARG(a && a->A && Whi && Wlo && (a->Y || a->Yt || Yhi || Ythi), "linear_tc: null pointer");
  OG_CHECK_ARG(a->rows > 0 && a->nout > 0 && a->batch > 0 && a->k1 > 0 && a->k2 >= 0, "linear_tc: bad sizes");
  OG_CHECK_ARG((Yhi == nullptr) == (Ylo == nullptr) && (Ythi == nullptr) == (Ytlo == nullptr), "linear_tc: hi/lo come in pairs");
  SplitOut so; so.Yhi = Yhi; so.Ylo = Ylo; so.Ythi = Ythi; so.Ytlo = Ytlo;
  return linear_tc_run(*a, Whi, Wlo, so, mode, (cudaStream_t)stream);
}

// one GEMM of the training step: tcgen05 3xTF32 when the shape is tileable (W is split on the fly into split_scratch, 2 x the
// W footprint in floats), the exact fp32 CUDA-core kernel otherwise
int64_t og_linear_auto_scratch_floats(const og_linear_args* a) {
  if (!a || a->nout <= 0 || a->ldw <= 0 || a->batch <= 0) return -1;
  const int64_t wf = (a->batch > 1 && a->strideW) ? (int64_t)(a->batch - 1) * a->strideW + (int64_t)a->nout * a->ldw : (int64_t)a->nout * a->ldw;
  return 2 * align_up(wf, 64);
}
int og_linear_auto_fwd(const og_linear_args* a, int precision, float* split_scratch, void* stream) {
  OG_CHECK_ARG(a && a->A && a->W && (a->Y || a->Yt), "linear_auto: null pointer");
  OG_CHECK_ARG(a->rows > 0 && a->nout > 0 && a->batch > 0 && a->k1 > 0 && a->k2 >= 0, "linear_auto: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != OG_PREC_FP32 && split_scratch && (!a->strideW || a->strideW % a->ldw == 0)) {
    const int64_t half = og_linear_auto_scratch_floats(a) / 2;
    const int64_t wf = (a->batch > 1 && a->strideW) ? (int64_t)(a->batch - 1) * a->strideW + (int64_t)a->nout * a->ldw : (int64_t)a->nout * a->ldw;
    float* hi = split_scratch; float* lo = split_scratch + half;
    TcLinearArgs t = to_tc_args(*a);
    if (linear_tc2_eligible(t, hi, lo, a->ldw) && (reinterpret_cast<uintptr_t>(a->W) & 15) == 0) {
      split_tf32_kernel<<<(unsigned)((wf + 255) / 256), 256, 0, st>>>(a->W, hi, lo, wf);
      OG_LAUNCH_CHECK("split_tf32_kernel");
      launch_counter()++;
      return linear_tc_run(*a, hi, lo, SplitOut(), 2, st);
    }
  }
  return linear_simt_launch(*a, st);
}

int og_split_tf32(const float* src, float* hi, float* lo, int64_t n, void* stream) {
  OG_CHECK_ARG(src && hi && lo && n > 0, "split_tf32: bad arguments");
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, hi, lo, n);
  OG_LAUNCH_CHECK("split_tf32_kernel");
  return OG_OK;
}

#ifdef OG_TRACE
int og_trace_read(long long* host_out) {     // debug build only
  OG_CUDA(cudaMemcpyFromSymbol(host_out, og_trace_buf, sizeof(long long) * 2 * 16 * 256));
  return OG_OK;
}
#endif

int og_attention_fwd(const float* q, int64_t ldq, int64_t strideq, const float* k, int64_t ldk, int64_t stridek,
                     const float* v, int64_t ldv, int64_t stridev, float* out, int64_t ldo, int64_t strideo,
                     int batch, int nq, int nk, int num_heads, int head_dim, int precision, void* stream) {
  OG_CHECK_ARG(q && k && v && out, "attention: null pointer");
  OG_CHECK_ARG(batch > 0 && nq > 0 && nk > 0 && num_heads > 0 && head_dim > 0, "attention: bad sizes");
  (void)precision;                                   // raw fp32 operands: always the exact kernel (see the header)
  AttnArgs a{q, ldq, strideq, k, ldk, stridek, v, ldv, stridev, out, ldo, strideo, batch, nq, nk, num_heads,
             (float)pow((double)head_dim, -0.5)};
  return attention_fp32(a, head_dim, (cudaStream_t)stream);
}

int og_attention_tc_fwd(const float* q, int64_t ldq, int64_t strideq, const float* khi, const float* klo, int64_t ldk,
                        const float* vthi, const float* vtlo, int64_t ldvt, float* out, int64_t ldo, int64_t strideo,
                        int batch, int nq, int nk, int num_heads, int head_dim, void* stream) {
  OG_CHECK_ARG(q && khi && klo && vthi && vtlo && out, "attention_tc: null pointer");
  OG_CHECK_ARG(batch > 0 && nq > 0 && nk > 0 && num_heads > 0, "attention_tc: bad sizes");
  if (!attention_tc_eligible(head_dim, ldq, ldk, ldvt, ldo))
    return fail(OG_EUNSUPPORTED, "attention_tc: head_dim in {32, 64} and 16-byte aligned rows required");
  TcAttnArgs a{q, ldq, strideq, out, ldo, strideo, batch, nq, nk, num_heads, num_heads * head_dim,
               (float)pow((double)head_dim, -0.5)};
  return attention_tc_launch(a, khi, klo, ldk, vthi, vtlo, ldvt, head_dim, (cudaStream_t)stream);
}

int64_t og_sinkhorn_workspace_bytes(int batch, int n, int m) { return sinkhorn_workspace_bytes(batch, n, m); }

int og_sinkhorn_fwd(const float* S, int64_t lds, int64_t strideS, const float* dustbin, int batch, int n, int m,
                    int iters, float reg, float* scores, void* workspace, int64_t workspace_bytes, void* stream) {
  OG_CHECK_ARG(S && dustbin && scores && workspace, "sinkhorn: null pointer");
  OG_CHECK_ARG(batch > 0 && n > 0 && m > 0 && iters >= 0 && reg > 0.f, "sinkhorn: bad sizes");
  return sinkhorn_launch(S, lds, strideS, dustbin, batch, n, m, iters, reg, scores, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

int64_t og_sinkhorn_hist_floats(int batch, int n, int m, int iters) {
  return (batch > 0 && n > 0 && m > 0 && iters >= 0) ? sinkhorn_hist_floats(batch, n, m, iters) : -1;
}

int og_sinkhorn_train_fwd(const float* S, int64_t lds, int64_t strideS, const float* dustbin, int batch, int n, int m,
                          int iters, float reg, float* scores, float* hist, void* workspace, int64_t workspace_bytes, void* stream) {
  OG_CHECK_ARG(S && dustbin && scores && workspace && hist, "sinkhorn_train_fwd: null pointer");
  OG_CHECK_ARG(batch > 0 && n > 0 && m > 0 && iters >= 0 && reg > 0.f, "sinkhorn_train_fwd: bad sizes");
  return sinkhorn_launch(S, lds, strideS, dustbin, batch, n, m, iters, reg, scores, workspace, workspace_bytes, (cudaStream_t)stream,
                         hist, hist + (int64_t)batch * iters * (n + 1));
}

int64_t og_sinkhorn_bwd_workspace_bytes(int batch, int n, int m, int iters) {
  return (batch > 0 && n > 0 && m > 0 && iters >= 0) ? sinkhorn_bwd_workspace_bytes(batch, n, m, iters) : -1;
}

int og_sinkhorn_bwd(const float* S, int64_t lds, int64_t strideS, const float* dustbin, int batch, int n, int m, int iters, float reg,
                    const float* hist, const float* dscores, float* dS_aug, float* ddustbin, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  OG_CHECK_ARG(S && dustbin && hist && dscores && dS_aug && ddustbin && workspace, "sinkhorn_bwd: null pointer");
  OG_CHECK_ARG(batch > 0 && n > 0 && m > 0 && iters >= 0 && reg > 0.f, "sinkhorn_bwd: bad sizes");
  return sinkhorn_bwd_launch(S, lds, strideS, dustbin, batch, n, m, iters, reg, hist, dscores, dS_aug, ddustbin, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

int64_t og_match_workspace_bytes(int batch, int n, int m) { return match_workspace_bytes(batch, n, m); }

int og_match_fwd(const float* scores, int batch, int n, int m, float threshold, int64_t* matches0, float* mscores0,
                 int64_t* matches1, float* mscores1, void* workspace, int64_t workspace_bytes, void* stream) {
  OG_CHECK_ARG(scores && workspace, "match: null pointer");
  OG_CHECK_ARG(batch > 0 && n > 0 && m > 0, "match: bad sizes");
  return match_launch(scores, batch, n, m, threshold, matches0, mscores0, matches1, mscores1, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}

int64_t og_gt_matches_workspace_bytes(int batch, int n, int m) {
  if (batch <= 0 || n <= 0 || m <= 0) return -1;
  return gt_matches_workspace_bytes(batch, n, m);
}

int og_gt_matches_fwd(const float* kpts0, const float* kpts1, int batch, int n, int m, const og_gt_transform* tf,
                      int64_t* gt_matches0, int64_t* gt_matches1, void* workspace, int64_t workspace_bytes, void* stream) {
  OG_CHECK_ARG(kpts0 && kpts1 && tf && gt_matches0 && gt_matches1 && workspace, "gt_matches: null pointer");
  OG_CHECK_ARG(batch > 0 && n > 0 && m > 0, "gt_matches: bad sizes (the reference returns (None, None) for an empty keypoint set)");
  OG_CHECK_ARG(tf->type == OG_GT_PERSPECTIVE || tf->type == OG_GT_3D_REPROJECTION, "gt_matches: unknown transformation type %d", tf->type);
  if (tf->type == OG_GT_PERSPECTIVE) {
    OG_CHECK_ARG(tf->H, "gt_matches: perspective transformation needs H");
  } else {
    OG_CHECK_ARG(tf->K0 && tf->K1 && tf->R && tf->T && tf->depth0 && tf->depth1, "gt_matches: 3d_reprojection needs K0, K1, R, T, depth0, depth1");
    OG_CHECK_ARG(!tf->depth_is_image || (tf->depth0_h > 0 && tf->depth0_w > 0 && tf->depth1_h > 0 && tf->depth1_w > 0),
                 "gt_matches: depth image sizes");
  }
  return gt_matches_launch(kpts0, kpts1, batch, n, m, *tf, gt_matches0, gt_matches1, workspace, workspace_bytes, (cudaStream_t)stream);
}

int og_collate_fwd(const float* lafs, const float* scores, const float* desc, const int* offsets, const int* select, int max_count,
                   const float* depth0, int depth0_h, int depth0_w, const float* depth1, int depth1_h, int depth1_w,
                   int batch, int target_keypoints, int descriptor_dim,
                   float* out_lafs0, float* out_lafs1, float* out_scores0, float* out_scores1, float* out_desc0, float* out_desc1,
                   float* out_depth0, float* out_depth1, void* stream) {
  OG_CHECK_ARG(lafs && scores && desc && offsets && out_lafs0 && out_lafs1 && out_scores0 && out_scores1 && out_desc0 && out_desc1,
               "collate: null pointer");
  OG_CHECK_ARG(batch > 0 && target_keypoints > 0 && descriptor_dim > 0 && max_count >= 0, "collate: bad sizes");
  OG_CHECK_ARG((depth0 == nullptr) == (out_depth0 == nullptr) && (depth1 == nullptr) == (out_depth1 == nullptr), "collate: depth in / out come together");
  OG_CHECK_ARG(!depth0 || (depth0_h > 0 && depth0_w > 0), "collate: depth0 size");
  OG_CHECK_ARG(!depth1 || (depth1_h > 0 && depth1_w > 0), "collate: depth1 size");
  CollateArgs a;
  a.lafs = lafs; a.scores = scores; a.desc = desc; a.offsets = offsets; a.select = select; a.depth0 = depth0; a.depth1 = depth1;
  a.B = batch; a.K = target_keypoints; a.D = descriptor_dim; a.h0 = depth0_h; a.w0 = depth0_w; a.h1 = depth1_h; a.w1 = depth1_w;
  a.out_lafs0 = out_lafs0; a.out_lafs1 = out_lafs1; a.out_scores0 = out_scores0; a.out_scores1 = out_scores1;
  a.out_desc0 = out_desc0; a.out_desc1 = out_desc1; a.out_depth0 = out_depth0; a.out_depth1 = out_depth1; a.sort_n = 1;
  return collate_launch(a, std::max(max_count, 1), (cudaStream_t)stream);
}

int64_t og_criterion_workspace_bytes(int batch) { return batch > 0 ? criterion_workspace_bytes(batch) : -1; }

int og_criterion_fwd(const float* scores, const int64_t* gt_matches0, const int64_t* gt_matches1, int batch, int n, int m,
                     float* loss, float* dscores, float grad_scale, void* workspace, int64_t workspace_bytes, void* stream) {
  OG_CHECK_ARG(scores && gt_matches0 && gt_matches1 && loss && workspace, "criterion: null pointer");
  OG_CHECK_ARG(batch > 0 && n > 0 && m > 0, "criterion: bad sizes");
  return criterion_launch(scores, gt_matches0, gt_matches1, batch, n, m, loss, dscores, grad_scale, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

// ---- training-step operators (row f1; csrc/train_ops.cuh) ----
int64_t og_train_workspace_floats(int cols) { return cols > 0 ? colreduce_workspace_floats(cols) + 2 * (int64_t)cols : -1; }

int og_transpose(const float* in, int64_t ld_in, int64_t stride_in, float* out, int64_t ld_out, int64_t stride_out,
                 int batch, int rows, int cols, int transpose, void* stream) {
  OG_CHECK_ARG(in && out, "transpose: null pointer");
  OG_CHECK_ARG(batch >= 0 && rows >= 0 && cols >= 0 && ld_in >= cols && ld_out >= (transpose ? rows : cols), "transpose: bad sizes");
  return transpose_launch(in, ld_in, stride_in, out, ld_out, stride_out, batch, rows, cols, transpose, (cudaStream_t)stream);
}

int og_colsum(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* z, int64_t ldz, int rows, int cols,
              float* out, float* workspace, void* stream) {
  OG_CHECK_ARG(x && out && workspace, "colsum: null pointer");
  OG_CHECK_ARG(rows >= 0 && cols > 0 && (!z || y), "colsum: bad arguments");
  ColReduceArgs a = {};
  a.x = x; a.ldx = ldx; a.y = y; a.ldy = ldy; a.z = z; a.ldz = ldz; a.rows = rows; a.cols = cols; a.partial = workspace; a.out0 = out;
  return colreduce_launch<0>(a, (cudaStream_t)stream);
}

int og_bn_train_fwd(const float* a_, int64_t lda, int rows, int cols, int relu, const float* gamma, const float* beta,
                    float eps, float momentum, float* y, int64_t ldy, float* save_mean, float* save_invstd,
                    float* running_mean, float* running_var, float* workspace, void* stream) {
  OG_CHECK_ARG(a_ && gamma && beta && y && save_mean && save_invstd && workspace, "bn_train_fwd: null pointer");
  OG_CHECK_ARG(rows > 0 && cols > 0, "bn_train_fwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  float* var = workspace;                                   // [cols]
  ColReduceArgs r = {};
  r.x = a_; r.ldx = lda; r.rows = rows; r.cols = cols; r.relu = relu; r.partial = workspace + 2 * (int64_t)cols; r.out0 = save_mean;
  int rc = colreduce_launch<1>(r, st);
  if (rc != OG_OK) return rc;
  r.mu = save_mean; r.out0 = var;
  if ((rc = colreduce_launch<2>(r, st)) != OG_OK) return rc;
  bn_finish_stats_kernel<<<cdiv(cols, 256), 256, 0, st>>>(save_mean, var, cols, rows, eps, momentum, save_invstd, running_mean, running_var);
  OG_LAUNCH_CHECK("bn_finish_stats_kernel");
  bn_apply_kernel<<<eltwise_grid((int64_t)rows * cols), 256, 0, st>>>(a_, lda, rows, cols, relu, save_mean, save_invstd, gamma, beta, y, ldy);
  OG_LAUNCH_CHECK("bn_apply_kernel");
  launch_counter() += 2;
  return OG_OK;
}

int og_bn_train_bwd(const float* dy, int64_t lddy, const float* a_, int64_t lda, int rows, int cols, int relu,
                    const float* gamma, const float* save_mean, const float* save_invstd,
                    float* da, int64_t ldda, float* dgamma, float* dbeta, float* workspace, void* stream) {
  OG_CHECK_ARG(dy && a_ && gamma && save_mean && save_invstd && da && dgamma && dbeta && workspace, "bn_train_bwd: null pointer");
  OG_CHECK_ARG(rows > 0 && cols > 0, "bn_train_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  ColReduceArgs r = {};
  r.x = dy; r.ldx = lddy; r.y = a_; r.ldy = lda; r.mu = save_mean; r.invstd = save_invstd; r.rows = rows; r.cols = cols; r.relu = relu;
  r.partial = workspace + 2 * (int64_t)cols; r.out0 = dbeta; r.out1 = dgamma;
  int rc = colreduce_launch<3>(r, st);
  if (rc != OG_OK) return rc;
  bn_bwd_apply_kernel<<<eltwise_grid((int64_t)rows * cols), 256, 0, st>>>(dy, lddy, a_, lda, rows, cols, relu, save_mean, save_invstd, gamma,
                                                                            dgamma, dbeta, da, ldda);
  OG_LAUNCH_CHECK("bn_bwd_apply_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_softmax_rows(float* S, int64_t ld, int64_t rows, int cols, void* stream) {
  OG_CHECK_ARG(S && rows >= 0 && cols > 0 && ld >= cols, "softmax_rows: bad arguments");
  if (rows == 0) return OG_OK;
  OG_CHECK_ARG((rows + 7) / 8 <= 0x7fffffffLL, "softmax_rows: too many rows");
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, ld, rows, cols);
  OG_LAUNCH_CHECK("softmax_rows_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_softmax_bwd_rows(const float* P, float* dP, int64_t ld, int64_t rows, int cols, float scale, void* stream) {
  OG_CHECK_ARG(P && dP && rows >= 0 && cols > 0 && ld >= cols, "softmax_bwd_rows: bad arguments");
  if (rows == 0) return OG_OK;
  OG_CHECK_ARG((rows + 7) / 8 <= 0x7fffffffLL, "softmax_bwd_rows: too many rows");
  softmax_bwd_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(P, dP, ld, rows, cols, scale);
  OG_LAUNCH_CHECK("softmax_bwd_rows_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_sum_batches(const float* part, int S, int rows, int cols, float* out, int64_t ld_out, int accumulate, void* stream) {
  OG_CHECK_ARG(part && out && S > 0 && rows > 0 && cols > 0 && ld_out >= cols, "sum_batches: bad arguments");
  sum_batches_kernel<<<eltwise_grid((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(part, S, rows, cols, out, ld_out, accumulate);
  OG_LAUNCH_CHECK("sum_batches_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_axpby(const float* x, const float* y, float a_, float b, float* out, int64_t n, void* stream) {
  OG_CHECK_ARG(x && out && n >= 0, "axpby: bad arguments");
  if (n == 0) return OG_OK;
  axpby_kernel<<<eltwise_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, a_, b, out, n);
  OG_LAUNCH_CHECK("axpby_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_mix_fwd(const float* g, const float* l, const float* mix, float* out, int64_t rows, int d, void* stream) {
  OG_CHECK_ARG(g && l && mix && out && rows > 0 && d > 0, "mix_fwd: bad arguments");
  mix_fwd_kernel<<<eltwise_grid(rows * d), 256, 0, (cudaStream_t)stream>>>(g, l, mix, out, rows, d);
  OG_LAUNCH_CHECK("mix_fwd_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_mix_bwd(const float* dm, const float* mix, float* dg, float* dl, int64_t rows, int d, void* stream) {
  OG_CHECK_ARG(dm && mix && (dg || dl) && rows > 0 && d > 0, "mix_bwd: bad arguments");
  mix_bwd_kernel<<<eltwise_grid(rows * d), 256, 0, (cudaStream_t)stream>>>(dm, mix, dg, dl, rows, d);
  OG_LAUNCH_CHECK("mix_bwd_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_mix_param_grad(const float* colsum, const float* mix, float* dmix, int d, void* stream) {
  OG_CHECK_ARG(colsum && mix && dmix && d > 0, "mix_param_grad: bad arguments");
  mix_param_grad_kernel<<<cdiv(d, 256), 256, 0, (cudaStream_t)stream>>>(colsum, mix, dmix, d);
  OG_LAUNCH_CHECK("mix_param_grad_kernel");
  launch_counter()++;
  return OG_OK;
}

int og_kenc_input(const float* kpts, const float* side, int rows, int side_info_size, float width, float height, float* out, void* stream) {
  OG_CHECK_ARG(kpts && out && rows > 0 && side_info_size >= 0 && (side_info_size == 0 || side), "kenc_input: bad arguments");
  kenc_input_kernel<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(kpts, side, rows, side_info_size, width - 1.f, height - 1.f, out);
  OG_LAUNCH_CHECK("kenc_input_kernel");
  launch_counter()++;
  return OG_OK;
}

// ---- SuperPoint front-end operators (row f4; csrc/superpoint.cuh) ----
int og_sp_im2col3x3(const float* x, int B, int H, int W, int C, float* out, void* stream) {
  OG_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && C > 0, "sp_im2col3x3: bad arguments");
  im2col3x3_kernel<<<sp_grid((int64_t)B * H * W * 9 * ((C % 4 == 0) ? C / 4 : C)), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, out);
  OG_LAUNCH_CHECK("im2col3x3_kernel");
  launch_counter()++;
  return OG_OK;
}
int og_sp_maxpool2x2(const float* x, int B, int H, int W, int C, float* out, void* stream) {
  OG_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "sp_maxpool2x2: bad arguments");
  maxpool2x2_kernel<<<sp_grid((int64_t)B * (H / 2) * (W / 2) * C), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, out);
  OG_LAUNCH_CHECK("maxpool2x2_kernel");
  launch_counter()++;
  return OG_OK;
}
int og_row_normalize(float* x, int64_t rows, int C, int mode, float eps, void* stream) {
  OG_CHECK_ARG(x && rows >= 0 && C > 0 && (mode == 0 || mode == 1), "row_normalize: bad arguments");
  if (rows == 0) return OG_OK;
  row_normalize_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, C, mode, eps);
  OG_LAUNCH_CHECK("row_normalize_kernel");
  launch_counter()++;
  return OG_OK;
}
int og_sp_heat_nms(const float* probs, int B, int Hc, int Wc, int nms_kernel, float threshold, int border, float* heat, void* stream) {
  OG_CHECK_ARG(probs && heat && B > 0 && Hc > 0 && Wc > 0 && nms_kernel > 0 && nms_kernel % 2 == 1 && border >= 0, "sp_heat_nms: bad arguments");
  sp_heat_nms_kernel<<<sp_grid((int64_t)B * Hc * Wc * 64), 256, 0, (cudaStream_t)stream>>>(probs, B, Hc, Wc, nms_kernel, threshold, border, heat);
  OG_LAUNCH_CHECK("sp_heat_nms_kernel");
  launch_counter()++;
  return OG_OK;
}
int og_sp_compact(const float* heat, int B, int HW, int cap, int* cand_idx, float* cand_score, int* count, void* stream) {
  OG_CHECK_ARG(heat && cand_idx && cand_score && count && B > 0 && HW > 0 && cap > 0, "sp_compact: bad arguments");
  sp_compact_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(heat, HW, cap, cand_idx, cand_score, count);
  OG_LAUNCH_CHECK("sp_compact_kernel");
  launch_counter()++;
  return OG_OK;
}
int og_sp_select(const int* cand_idx, const float* cand_score, const int* count, const int* n_out, const int* mode, int B, int cap, int W,
                 int out_cap, int max_count, float* kpts, float* scores, void* stream) {
  OG_CHECK_ARG(cand_idx && cand_score && count && n_out && mode && kpts && scores, "sp_select: null pointer");
  OG_CHECK_ARG(B > 0 && cap > 0 && W > 0 && out_cap > 0 && max_count >= 0, "sp_select: bad sizes");
  if (max_count > SP_MAX_CAND) return fail(OG_EUNSUPPORTED, "sp_select: %d candidates in one image exceed the sort capacity %d", max_count, SP_MAX_CAND);
  int n2 = 1;
  while (n2 < max_count) n2 <<= 1;
  const int smem = 2 * n2 * 4;
  static DeviceFlags attr_set;
  if (attr_set.once()) OG_CUDA(cudaFuncSetAttribute(sp_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SP_MAX_CAND * 4));
  sp_select_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(cand_idx, cand_score, count, n_out, mode, cap, W, out_cap, kpts, scores);
  OG_LAUNCH_CHECK("sp_select_kernel");
  launch_counter()++;
  return OG_OK;
}
int og_sp_sample_desc(const float* coarse, int B, int Hc, int Wc, int D, const float* kpts, const int* n_out, int out_cap, int max_n, int cell,
                      float* desc, void* stream) {
  OG_CHECK_ARG(coarse && kpts && n_out && desc && B > 0 && Hc > 0 && Wc > 0 && D > 0 && out_cap > 0 && cell > 0, "sp_sample_desc: bad arguments");
  if (max_n <= 0) return OG_OK;
  sp_sample_desc_kernel<<<dim3(cdiv(max_n, 8), B), 256, 0, (cudaStream_t)stream>>>(coarse, Hc, Wc, D, kpts, n_out, out_cap, cell, desc);
  OG_LAUNCH_CHECK("sp_sample_desc_kernel");
  launch_counter()++;
  return OG_OK;
}

static int forward_impl(const og_config* cfg, const float* Wp, const float* Whi, const float* Wlo, const __half* W16h,
                        const __half* W16l, const float* meta16, int B, int n, int m,
                         const float* kpts0,
                         const float* kpts1, const float* side0, const float* side1, const float* desc0,
                         const float* desc1, const float* img_wh, float* ctx0, float* ctx1, float* scores,
                         int64_t* matches0, float* mscores0, int64_t* matches1, float* mscores1, void* workspace,
                         int64_t workspace_bytes, void* stream_) {
  int rc = check_config(cfg);
  if (rc != OG_OK) return rc;
  OG_CHECK_ARG(Wp && kpts0 && kpts1 && desc0 && desc1 && img_wh && scores && workspace, "forward: null pointer");
  OG_CHECK_ARG(cfg->side_info_size == 0 || (side0 && side1), "forward: side info missing");
  OG_CHECK_ARG(B > 0 && n > 0 && m > 0, "forward: batch, n, m must be positive");
  OG_CHECK_ARG(cfg->precision == OG_PREC_FP32 || (Whi && Wlo), "forward: the tensor-core modes need packed_hi / packed_lo");
  OG_CHECK_ARG(cfg->precision != OG_PREC_FP16X3 || (W16h && W16l && meta16), "forward: OG_PREC_FP16X3 needs the fp16 weight split (og_pack_f16)");
  const bool tcp = (cfg->precision != OG_PREC_FP32) && cfg->descriptor_dim >= 32;   // K >= 32 for the tcgen05 tiles
  // fp16 hi/lo form of the GNN (projections, attention, message MLP): head_dim 64, d a multiple of 64; everything else
  // (Dh = 32, the final projection, the score GEMM) runs the tf32 hi/lo form - same accuracy class
  const bool f16p = tcp && cfg->precision == OG_PREC_FP16X3 && cfg->descriptor_dim % 64 == 0 &&
                    cfg->descriptor_dim / cfg->num_heads == 64;
  auto WH = [&](int64_t off) { return tcp ? Whi + off : nullptr; };
  auto WL = [&](int64_t off) { return tcp ? Wlo + off : nullptr; };
  OG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "forward: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream_;
  const int prec = cfg->precision == OG_PREC_FP16X3 ? OG_PREC_TF32X3 : cfg->precision;     // what the non-f16 launches run as
  const Layout L = make_layout(cfg);
  Workspace w;
  rc = plan_workspace(cfg, B, n, m, workspace, &w);
  if (rc != OG_OK) return rc;
  if (workspace_bytes < w.total) return fail(OG_EWORKSPACE, "forward: workspace %lld < %lld bytes",
                                             (long long)workspace_bytes, (long long)w.total);
  launch_counter() = 0;
  const int d = cfg->descriptor_dim, H = cfg->num_heads, dh = d / H, S = cfg->side_info_size;
  const int R0 = B * n, R1 = B * m, R = R0 + R1;
  float* x0 = w.x; float* x1 = w.x + (int64_t)R0 * d;

  // ---- keypoint encoder (positional_encoding.py:16-19) + descriptors (superglue.py:52-55) ----
  kenc_input_kernel<<<cdiv(R0, 256), 256, 0, st>>>(kpts0, side0, R0, S, img_wh[0] - 1.f, img_wh[1] - 1.f, w.in0);
  OG_LAUNCH_CHECK("kenc_input_kernel");
  kenc_input_kernel<<<cdiv(R1, 256), 256, 0, st>>>(kpts1, side1, R1, S, img_wh[2] - 1.f, img_wh[3] - 1.f,
                                                    w.in0 + (int64_t)R0 * (2 + S));
  OG_LAUNCH_CHECK("kenc_input_kernel");
  launch_counter() += 2;
  {
    const float* cur = w.in0;
    float* bufs[2] = {w.h0, w.h1};
    const int nl = (int)L.kenc_sizes.size() - 1;
    for (int i = 0; i < nl; ++i) {
      const int kin = L.kenc_sizes[i], kout = L.kenc_sizes[i + 1];
      if (i < nl - 1) {
        og_linear_args a = lin(cur, kin, kin, Wp + L.kenc_w[i], Wp + L.kenc_b[i], R, kout, bufs[i & 1], kout);
        a.relu = 1;
        if ((rc = linear_dispatch(a, OG_PREC_FP32, st)) != OG_OK) return rc;
        cur = bufs[i & 1];
      } else {                        // last layer: + local descriptors, per image (separate user tensors)
        og_linear_args a = lin(cur, kin, kin, Wp + L.kenc_w[i], Wp + L.kenc_b[i], R0, kout, x0, d);
        if (!cfg->no_descriptors) { a.R = desc0; a.ldr = d; }          // superglue.py:45-55
        if ((rc = linear_dispatch(a, OG_PREC_FP32, st)) != OG_OK) return rc;
        og_linear_args b = lin(cur + (int64_t)R0 * kin, kin, kin, Wp + L.kenc_w[i], Wp + L.kenc_b[i], R1, kout, x1, d);
        if (!cfg->no_descriptors) { b.R = desc1; b.ldr = d; }
        if ((rc = linear_dispatch(b, OG_PREC_FP32, st)) != OG_OK) return rc;
      }
    }
  }

  // ---- attentional GNN (attention_gnn.py:58-93) ----
  auto attend = [&](int qrow0, int nq, int krow0, int nk, int batch) -> int {
    // q from qkv[:, 0:d] of rows qrow0.., k/v from qkv[:, d:3d] of rows krow0..; out -> o rows qrow0..
    AttnArgs a{w.qkv + (int64_t)qrow0 * 3 * d, 3 * d, (int64_t)nq * 3 * d,
               w.qkv + (int64_t)krow0 * 3 * d + d, 3 * d, (int64_t)nk * 3 * d,
               w.qkv + (int64_t)krow0 * 3 * d + 2 * d, 3 * d, (int64_t)nk * 3 * d,
               w.o + (int64_t)qrow0 * d, d, (int64_t)nq * d, batch, nq, nk, H, (float)pow((double)dh, -0.5)};
    return attention_fp32(a, dh, st);
  };
  auto mlp = [&](int l, int row0, int rows) -> int {
    // x <- x + W2 . relu(W1 . [x ; o] + b1) + b2     (out_proj and BN folded into W1 / W2 on the host)
    float* xr = w.x + (int64_t)row0 * d;
    og_linear_args a = lin(xr, d, d, Wp + L.fc1_w[l], Wp + L.fc1_b[l], rows, 2 * d, w.hid + (int64_t)row0 * 2 * d, 2 * d);
    a.A2 = w.o + (int64_t)row0 * d; a.lda2 = d; a.k2 = d; a.ldw = 2 * d; a.relu = 1;
    int r = linear_dispatch(a, prec, st, WH(L.fc1_w[l]), WL(L.fc1_w[l]));
    if (r != OG_OK) return r;
    og_linear_args c2 = lin(w.hid + (int64_t)row0 * 2 * d, 2 * d, 2 * d, Wp + L.fc2_w[l], Wp + L.fc2_b[l], rows, d, xr, d);
    c2.R = xr; c2.ldr = d;
    return linear_dispatch(c2, prec, st, WH(L.fc2_w[l]), WL(L.fc2_w[l]));
  };
  auto project = [&](int l, int row0, int rows, int wrow0, int nout) -> int {
    // qkv[rows, wrow0 : wrow0 + nout] = x[rows] . Wqkv[wrow0 : wrow0 + nout]^T + b
    og_linear_args a = lin(w.x + (int64_t)row0 * d, d, d, Wp + L.qkv_w[l] + (int64_t)wrow0 * d, Wp + L.qkv_b[l] + wrow0,
                           rows, nout, w.qkv + (int64_t)row0 * 3 * d + wrow0, 3 * d);
    return linear_dispatch(a, prec, st, WH(L.qkv_w[l] + (int64_t)wrow0 * d), WL(L.qkv_w[l] + (int64_t)wrow0 * d));
  };
  // Tensor-core attention path (OG_PREC_TF32X3, Dh in {32, 64}): the projection GEMM writes Q as fp32, K split
  // hi/lo keypoint-major and V split hi/lo channel-major (= the reference's own [B, d, M] layout), which are
  // exactly the operand layouts csrc/attention_tc.cuh stages with TMA.
  const bool tca_ok = tcp && (dh == 32 || dh == 64);
  auto project_tc = [&](int l, int qrow0, int nq_rows, int srow0, int ns, int sbatch) -> int {
    // q rows [qrow0, +nq_rows);  k / v from source rows [srow0, +sbatch*ns) (sbatch sequences of ns keypoints)
    int r;
    og_linear_args aq = lin(w.x + (int64_t)qrow0 * d, d, d, Wp + L.qkv_w[l], Wp + L.qkv_b[l], nq_rows, d,
                            w.q + (int64_t)qrow0 * d, d);
    if ((r = linear_dispatch(aq, prec, st, WH(L.qkv_w[l]), WL(L.qkv_w[l]))) != OG_OK) return r;
    og_linear_args ak = lin(w.x + (int64_t)srow0 * d, d, d, Wp + L.qkv_w[l] + (int64_t)d * d, Wp + L.qkv_b[l] + d,
                            sbatch * ns, d, nullptr, d);
    SplitOut sk; sk.Yhi = w.khi + (int64_t)srow0 * d; sk.Ylo = w.klo + (int64_t)srow0 * d;
    if ((r = linear_dispatch(ak, prec, st, WH(L.qkv_w[l] + (int64_t)d * d), WL(L.qkv_w[l] + (int64_t)d * d), sk)) != OG_OK) return r;
    const int64_t ldv = (srow0 == 0) ? w.ldn : w.ldm;
    const int64_t voff = (srow0 == 0) ? 0 : (int64_t)B * d * w.ldn;
    og_linear_args av = lin(w.x + (int64_t)srow0 * d, d, d, Wp + L.qkv_w[l] + 2 * (int64_t)d * d, Wp + L.qkv_b[l] + 2 * d,
                            ns, d, nullptr, d);
    av.batch = sbatch; av.strideA = (int64_t)ns * d; av.ldyt = ldv; av.strideYt = (int64_t)d * ldv;
    SplitOut sv; sv.Ythi = w.vthi + voff; sv.Ytlo = w.vtlo + voff;
    return linear_dispatch(av, prec, st, WH(L.qkv_w[l] + 2 * (int64_t)d * d), WL(L.qkv_w[l] + 2 * (int64_t)d * d), sv);
  };
  auto attend_tc = [&](int qrow0, int nq, int krow0, int nk, int batch) -> int {
    const int64_t ldv = (krow0 == 0) ? w.ldn : w.ldm;
    const int64_t voff = (krow0 == 0) ? 0 : (int64_t)B * d * w.ldn;
    TcAttnArgs a{w.q + (int64_t)qrow0 * d, d, (int64_t)nq * d, w.o + (int64_t)qrow0 * d, d, (int64_t)nq * d,
                 batch, nq, nk, H, d, (float)pow((double)dh, -0.5)};
    return attention_tc_launch(a, w.khi + (int64_t)krow0 * d, w.klo + (int64_t)krow0 * d, d, w.vthi + voff, w.vtlo + voff,
                               ldv, dh, st);
  };
  // ---- fp16 hi/lo form of the GNN (OG_PREC_FP16X3; csrc/linear_f16.cuh, csrc/attention_f16.cuh) ----
  // Every activation tensor has a device slot with its tracked amax (fp32 tensors) or the scale it was written with (fp16
  // K / V^T); a consumer derives its operand scale from the producers' slots, so no scale ever passes through the host.
  int nslot = 0;
  auto new_slot = [&]() -> float* { return w.slots + (nslot < w.nslots - 1 ? nslot++ : w.nslots - 1); };
  float *sx0 = nullptr, *sx1 = nullptr;                      // amax slots of the current x0 / x1
  __half* kh16 = reinterpret_cast<__half*>(w.khi); __half* kl16 = reinterpret_cast<__half*>(w.klo);
  __half* vth16 = reinterpret_cast<__half*>(w.vthi); __half* vtl16 = reinterpret_cast<__half*>(w.vtlo);
  if (f16p) {
    OG_CUDA(cudaMemsetAsync(w.slots, 0, (size_t)w.nslots * 4, st));
    sx0 = sx1 = new_slot();
    const int64_t nx = (int64_t)R * d;
    amax_kernel<<<(unsigned)std::min<int64_t>((nx + 2047) / 2048, 1184), 256, 0, st>>>(w.x, nx, sx0);
    OG_LAUNCH_CHECK("amax_kernel");
    launch_counter()++;
  }
  auto M16 = [&](int l, int t) { return meta16 + ((int64_t)l * 5 + t) * 4; };
  auto gemm16 = [&](const float* A, int64_t lda, int k1, const float* A2, int64_t lda2, int k2, int64_t woff, const float* bias,
                    int rows, int nout, const float* am0, const float* am1, const float* am2, const float* meta) {
    F16LinearArgs g;
    memset(&g, 0, sizeof(g));
    g.A = A; g.lda = lda; g.k1 = k1; g.A2 = A2; g.lda2 = lda2; g.k2 = k2; g.bias = bias; g.rows = rows; g.nout = nout; g.batch = 1;
    g.alpha = 1.f; g.amax_in[0] = am0; g.amax_in[1] = am1; g.amax_in[2] = am2; g.w_meta = meta;
    (void)woff;
    return g;
  };
  auto run16 = [&](const F16LinearArgs& g, int64_t woff) -> int {
    const int K = g.k1 + g.k2;
    if (!linear_f16_eligible(g, W16h + woff, W16l + woff, K)) return fail(OG_EUNSUPPORTED, "forward: a layer is not tileable by the fp16 GEMM");
    return linear_f16_launch(g, W16h + woff, W16l + woff, K, g.nout, st);
  };
  struct SeqSlots { float *q, *k, *v, *o; };
  // Q / K / V projections.  d % 128 == 0 (every shipped config): the projections that share their input run as ONE launch over the
  // stacked weights (linear_f16.cuh OUTK 4) - Q | K | V in a self layer (same keypoints), K | V in a cross layer (+ Q on its own).
  const int fuse_qkv = fuse_qkv_mode();
  auto project_f16 = [&](int l, int qrow0, int nq_rows, int srow0, int ns, int sbatch, float* aq0, float* aq1, float* as0, float* as1,
                         SeqSlots& ss) -> int {
    int r;
    ss.q = new_slot(); ss.k = new_slot(); ss.v = new_slot(); ss.o = new_slot();
    const int64_t ldv = (srow0 == 0) ? w.ldn : w.ldm;
    const int64_t voff = (srow0 == 0) ? 0 : (int64_t)B * d * w.ldn;
    const bool fuse = fuse_qkv && d % tcf::BN == 0;
    const bool same_src = qrow0 == srow0 && nq_rows == sbatch * ns;
    if (!(fuse && same_src)) {
      F16LinearArgs gq = gemm16(w.x + (int64_t)qrow0 * d, d, d, nullptr, 0, 0, L.qkv_w[l], Wp + L.qkv_b[l], nq_rows, d, aq0, aq1, nullptr, M16(l, 0));
      gq.Y = w.q + (int64_t)qrow0 * d; gq.ldy = d; gq.amax_out = ss.q;
      if ((r = run16(gq, L.qkv_w[l])) != OG_OK) return r;
    }
    if (fuse) {
      const int k0 = same_src ? 0 : 1, nk_ = 3 - k0;
      F16LinearArgs g = gemm16(w.x + (int64_t)srow0 * d, d, d, nullptr, 0, 0, 0, Wp + L.qkv_b[l] + k0 * d, ns, nk_ * d, as0, as1, nullptr, M16(l, k0));
      g.batch = sbatch; g.strideA = (int64_t)ns * d;
      g.nkinds = nk_; g.kind0 = k0; g.kind_cols = d;
      g.ldy = d; g.strideY = (int64_t)ns * d;
      if (k0 == 0) { g.Y = w.q + (int64_t)srow0 * d; g.amax_out = ss.q; }
      g.Yh = kh16 + (int64_t)srow0 * d; g.Yl = kl16 + (int64_t)srow0 * d; g.scale_out = ss.k;
      g.Yth = vth16 + voff; g.Ytl = vtl16 + voff; g.ldyt = ldv; g.strideYt = (int64_t)d * ldv; g.scale_out_v = ss.v;
      return run16(g, L.qkv_w[l] + (int64_t)k0 * d * d);
    }
    F16LinearArgs gk = gemm16(w.x + (int64_t)srow0 * d, d, d, nullptr, 0, 0, 0, Wp + L.qkv_b[l] + d, sbatch * ns, d, as0, as1, nullptr, M16(l, 1));
    gk.Yh = kh16 + (int64_t)srow0 * d; gk.Yl = kl16 + (int64_t)srow0 * d; gk.ldy = d; gk.scale_out = ss.k;
    if ((r = run16(gk, L.qkv_w[l] + (int64_t)d * d)) != OG_OK) return r;
    F16LinearArgs gv = gemm16(w.x + (int64_t)srow0 * d, d, d, nullptr, 0, 0, 0, Wp + L.qkv_b[l] + 2 * d, ns, d, as0, as1, nullptr, M16(l, 2));
    gv.batch = sbatch; gv.strideA = (int64_t)ns * d; gv.Yth = vth16 + voff; gv.Ytl = vtl16 + voff; gv.ldyt = ldv; gv.strideYt = (int64_t)d * ldv;
    gv.scale_out = ss.v;
    return run16(gv, L.qkv_w[l] + 2 * (int64_t)d * d);
  };
  auto attend_f16 = [&](int qrow0, int nq, int krow0, int nk, int batch, const SeqSlots& ss) -> int {
    const int64_t ldv = (krow0 == 0) ? w.ldn : w.ldm;
    const int64_t voff = (krow0 == 0) ? 0 : (int64_t)B * d * w.ldn;
    TcAttnArgs a{w.q + (int64_t)qrow0 * d, d, (int64_t)nq * d, w.o + (int64_t)qrow0 * d, d, (int64_t)nq * d,
                 batch, nq, nk, H, d, (float)pow((double)dh, -0.5)};
    F16AttnScales sc{ss.q, ss.k, ss.v, ss.o, 0};
    return attention_f16_dispatch(a, sc, kh16 + (int64_t)krow0 * d, kl16 + (int64_t)krow0 * d, d, vth16 + voff, vtl16 + voff, ldv, dh, st);
  };
  auto mlp_f16 = [&](int l, int row0, int rows, float* ax0, float* ax1, float* ao, float** sx_new) -> int {
    float* sh = new_slot();
    float* sn = new_slot();
    float* xr = w.x + (int64_t)row0 * d;
    F16LinearArgs g1 = gemm16(xr, d, d, w.o + (int64_t)row0 * d, d, d, 0, Wp + L.fc1_b[l], rows, 2 * d, ax0, ax1, ao, M16(l, 3));
    g1.relu = 1; g1.Y = w.hid + (int64_t)row0 * 2 * d; g1.ldy = 2 * d; g1.amax_out = sh;
    int r = run16(g1, L.fc1_w[l]);
    if (r != OG_OK) return r;
    F16LinearArgs g2 = gemm16(w.hid + (int64_t)row0 * 2 * d, 2 * d, 2 * d, nullptr, 0, 0, 0, Wp + L.fc2_b[l], rows, d, sh, nullptr, nullptr, M16(l, 4));
    g2.R = xr; g2.ldr = d; g2.Y = xr; g2.ldy = d; g2.amax_out = sn;
    *sx_new = sn;
    return run16(g2, L.fc2_w[l]);
  };
  for (int l = 0; l < cfg->num_layers; ++l) {
    if (f16p) {
      SeqSlots ss;
      float* sn = nullptr;
      if (l % 2 == 0) {                                    // self
        if (n == m) {
          if ((rc = project_f16(l, 0, R, 0, n, 2 * B, sx0, sx1, sx0, sx1, ss)) != OG_OK) return rc;
          if ((rc = attend_f16(0, n, 0, n, 2 * B, ss)) != OG_OK) return rc;
          if ((rc = mlp_f16(l, 0, R, sx0, sx1, ss.o, &sn)) != OG_OK) return rc;
          sx0 = sx1 = sn;
        } else {
          SeqSlots s1;
          if ((rc = project_f16(l, 0, R0, 0, n, B, sx0, nullptr, sx0, nullptr, ss)) != OG_OK) return rc;
          if ((rc = attend_f16(0, n, 0, n, B, ss)) != OG_OK) return rc;
          if ((rc = project_f16(l, R0, R1, R0, m, B, sx1, nullptr, sx1, nullptr, s1)) != OG_OK) return rc;
          if ((rc = attend_f16(R0, m, R0, m, B, s1)) != OG_OK) return rc;
          float *sa = nullptr, *sb = nullptr;
          if ((rc = mlp_f16(l, 0, R0, sx0, nullptr, ss.o, &sa)) != OG_OK) return rc;
          if ((rc = mlp_f16(l, R0, R1, sx1, nullptr, s1.o, &sb)) != OG_OK) return rc;
          sx0 = sa; sx1 = sb;
        }
      } else {                                             // cross: SEQUENTIAL (attention_gnn.py:74-77)
        if ((rc = project_f16(l, 0, R0, R0, m, B, sx0, nullptr, sx1, nullptr, ss)) != OG_OK) return rc;
        if ((rc = attend_f16(0, n, R0, m, B, ss)) != OG_OK) return rc;
        if ((rc = mlp_f16(l, 0, R0, sx0, nullptr, ss.o, &sn)) != OG_OK) return rc;
        sx0 = sn;
        if ((rc = project_f16(l, R0, R1, 0, n, B, sx1, nullptr, sx0, nullptr, ss)) != OG_OK) return rc;     // k, v of the UPDATED image 0
        if ((rc = attend_f16(R0, m, 0, n, B, ss)) != OG_OK) return rc;
        if ((rc = mlp_f16(l, R0, R1, sx1, nullptr, ss.o, &sn)) != OG_OK) return rc;
        sx1 = sn;
      }
      continue;
    }
    if (tca_ok) {
      if (l % 2 == 0) {                                    // self
        if (n == m) {
          if ((rc = project_tc(l, 0, R, 0, n, 2 * B)) != OG_OK) return rc;
          if ((rc = attend_tc(0, n, 0, n, 2 * B)) != OG_OK) return rc;
        } else {
          if ((rc = project_tc(l, 0, R0, 0, n, B)) != OG_OK) return rc;
          if ((rc = attend_tc(0, n, 0, n, B)) != OG_OK) return rc;
          if ((rc = project_tc(l, R0, R1, R0, m, B)) != OG_OK) return rc;
          if ((rc = attend_tc(R0, m, R0, m, B)) != OG_OK) return rc;
        }
        if ((rc = mlp(l, 0, R)) != OG_OK) return rc;
      } else {                                             // cross: SEQUENTIAL (attention_gnn.py:74-77)
        if ((rc = project_tc(l, 0, R0, R0, m, B)) != OG_OK) return rc;
        if ((rc = attend_tc(0, n, R0, m, B)) != OG_OK) return rc;
        if ((rc = mlp(l, 0, R0)) != OG_OK) return rc;
        if ((rc = project_tc(l, R0, R1, 0, n, B)) != OG_OK) return rc;     // k, v of the UPDATED image 0
        if ((rc = attend_tc(R0, m, 0, n, B)) != OG_OK) return rc;
        if ((rc = mlp(l, R0, R1)) != OG_OK) return rc;
      }
      continue;
    }
    if (l % 2 == 0) {                                      // self: both images, shared weights, independent
      if ((rc = project(l, 0, R, 0, 3 * d)) != OG_OK) return rc;
      if (n == m) {
        if ((rc = attend(0, n, 0, n, 2 * B)) != OG_OK) return rc;
      } else {
        if ((rc = attend(0, n, 0, n, B)) != OG_OK) return rc;
        if ((rc = attend(R0, m, R0, m, B)) != OG_OK) return rc;
      }
      if ((rc = mlp(l, 0, R)) != OG_OK) return rc;
    } else {                                               // cross: SEQUENTIAL (attention_gnn.py:74-77)
      if ((rc = project(l, 0, R0, 0, d)) != OG_OK) return rc;            // q of image 0
      if ((rc = project(l, R0, R1, d, 2 * d)) != OG_OK) return rc;       // k, v of image 1
      if ((rc = attend(0, n, R0, m, B)) != OG_OK) return rc;
      if ((rc = mlp(l, 0, R0)) != OG_OK) return rc;
      if ((rc = project(l, R0, R1, 0, d)) != OG_OK) return rc;           // q of image 1
      if ((rc = project(l, 0, R0, d, 2 * d)) != OG_OK) return rc;        // k, v of the UPDATED image 0
      if ((rc = attend(R0, m, 0, n, B)) != OG_OK) return rc;
      if ((rc = mlp(l, R0, R1)) != OG_OK) return rc;
    }
  }

  // ---- final projection + residual mix (superglue.py:58-62); channel-first context descriptors ----
  for (int img = 0; img < 2; ++img) {
    const int nn = img ? m : n;
    float* xr = img ? x1 : x0;
    float* gr = w.g + (img ? (int64_t)R0 * d : 0);
    og_linear_args a = lin(xr, d, d, Wp + L.proj_w, Wp + L.proj_b, nn, d, gr, d);
    a.batch = B; a.strideA = (int64_t)nn * d; a.strideY = (int64_t)nn * d;
    a.R = img ? desc1 : desc0; a.ldr = d; a.strideR = (int64_t)nn * d; a.rscale = Wp + L.proj_rmix;
    float* ctx = img ? ctx1 : ctx0;
    if (ctx) { a.Yt = ctx; a.ldyt = nn; a.strideYt = (int64_t)d * nn; }
    SplitOut so;
    if (tcp && img == 1) { so.Yhi = w.ghi; so.Ylo = w.glo; }      // image-1 descriptors are the score GEMM's B operand
    if ((rc = linear_dispatch(a, prec, st, WH(L.proj_w), WL(L.proj_w), so)) != OG_OK) return rc;
  }
  // ---- score matrix (superglue.py:64,80-86): S = g0^T g1 * d^-0.5, written with padded rows ----
  {
    og_linear_args a = lin(w.g, d, d, w.g + (int64_t)R0 * d, nullptr, n, m, w.sbuf, w.lds);
    a.batch = B; a.strideA = (int64_t)n * d; a.strideW = (int64_t)m * d; a.strideY = (int64_t)n * w.lds;
    a.alpha = (float)pow((double)d, -0.5);
    if ((rc = linear_dispatch(a, prec, st, tcp ? w.ghi : nullptr, tcp ? w.glo : nullptr)) != OG_OK) return rc;
  }
  // ---- optimal transport (superglue.py:88-111) + matches (matching_module.py:174-187) ----
  rc = sinkhorn_launch(w.sbuf, w.lds, (int64_t)n * w.lds, Wp + L.dustbin, B, n, m, cfg->sinkhorn_iters,
                       cfg->sinkhorn_reg, scores, w.sink, w.sink_bytes, st);
  if (rc != OG_OK) return rc;
  if (matches0 || mscores0 || matches1 || mscores1) {
    rc = match_launch(scores, B, n, m, cfg->match_threshold, matches0, mscores0, matches1, mscores1, w.match,
                      w.match_bytes, st);
    if (rc != OG_OK) return rc;
  }
  return OG_OK;
}

int og_superglue_forward(const og_config* cfg, const float* Wp, const float* Whi, const float* Wlo, int B, int n, int m,
                         const float* kpts0, const float* kpts1, const float* side0, const float* side1, const float* desc0,
                         const float* desc1, const float* img_wh, float* ctx0, float* ctx1, float* scores,
                         int64_t* matches0, float* mscores0, int64_t* matches1, float* mscores1, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  if (cfg && cfg->precision == OG_PREC_FP16X3) return fail(OG_EINVAL, "forward: OG_PREC_FP16X3 takes og_superglue_forward_f16");
  return forward_impl(cfg, Wp, Whi, Wlo, nullptr, nullptr, nullptr, B, n, m, kpts0, kpts1, side0, side1, desc0, desc1, img_wh, ctx0, ctx1,
                      scores, matches0, mscores0, matches1, mscores1, workspace, workspace_bytes, stream);
}

int og_superglue_forward_f16(const og_config* cfg, const float* Wp, const float* Whi, const float* Wlo, const void* W16h,
                             const void* W16l, const float* meta16, int B, int n, int m,
                             const float* kpts0, const float* kpts1, const float* side0, const float* side1, const float* desc0,
                             const float* desc1, const float* img_wh, float* ctx0, float* ctx1, float* scores,
                             int64_t* matches0, float* mscores0, int64_t* matches1, float* mscores1, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  return forward_impl(cfg, Wp, Whi, Wlo, static_cast<const __half*>(W16h), static_cast<const __half*>(W16l), meta16, B, n, m, kpts0, kpts1,
                      side0, side1, desc0, desc1, img_wh, ctx0, ctx1, scores, matches0, mscores0, matches1, mscores1, workspace,
                      workspace_bytes, stream);
}

int64_t og_f16_meta_floats(const og_config* cfg) {
  if (check_config(cfg) != OG_OK) return -1;
  return (int64_t)cfg->num_layers * 5 * 4;
}

int og_weight_split_f16(const float* w, const float* bias, int rows, int cols, void* hi16, void* lo16, float* meta, void* stream) {
  OG_CHECK_ARG(w && hi16 && lo16 && meta && rows > 0 && cols > 0, "weight_split_f16: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  OG_CUDA(cudaMemsetAsync(meta, 0, 4 * sizeof(float), st));
  weight_meta_kernel<<<cdiv(rows, 8), 256, 0, st>>>(w, bias, rows, cols, meta);
  OG_LAUNCH_CHECK("weight_meta_kernel");
  const int64_t nel = (int64_t)rows * cols;
  split_f16_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(w, static_cast<__half*>(hi16), static_cast<__half*>(lo16), nel, meta);
  OG_LAUNCH_CHECK("split_f16_kernel");
  finish_meta_kernel<<<1, 1, 0, st>>>(meta);
  OG_LAUNCH_CHECK("finish_meta_kernel");
  return OG_OK;
}

int og_pack_f16(const og_config* cfg, const float* Wp, void* hi16, void* lo16, float* meta, void* stream) {
  int rc = check_config(cfg);
  if (rc != OG_OK) return rc;
  OG_CHECK_ARG(Wp && hi16 && lo16 && meta, "pack_f16: null pointer");
  const Layout L = make_layout(cfg);
  const int64_t d = cfg->descriptor_dim;
  __half* h = static_cast<__half*>(hi16); __half* lo = static_cast<__half*>(lo16);
  for (int l = 0; l < cfg->num_layers; ++l) {
    struct T { int64_t woff, boff; int rows, cols; } ts[5] = {
      {L.qkv_w[l], L.qkv_b[l], (int)d, (int)d}, {L.qkv_w[l] + d * d, L.qkv_b[l] + d, (int)d, (int)d},
      {L.qkv_w[l] + 2 * d * d, L.qkv_b[l] + 2 * d, (int)d, (int)d},
      {L.fc1_w[l], L.fc1_b[l], (int)(2 * d), (int)(2 * d)}, {L.fc2_w[l], L.fc2_b[l], (int)d, (int)(2 * d)}};
    for (int t = 0; t < 5; ++t)
      if ((rc = og_weight_split_f16(Wp + ts[t].woff, Wp + ts[t].boff, ts[t].rows, ts[t].cols, h + ts[t].woff, lo + ts[t].woff,
                                    meta + ((int64_t)l * 5 + t) * 4, stream)) != OG_OK) return rc;
  }
  return OG_OK;
}

int og_amax(const float* x, int64_t n, float* slot, void* stream) {
  OG_CHECK_ARG(x && slot && n > 0, "amax: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  OG_CUDA(cudaMemsetAsync(slot, 0, sizeof(float), st));
  amax_kernel<<<(unsigned)std::min<int64_t>((n + 2047) / 2048, 1184), 256, 0, st>>>(x, n, slot);
  OG_LAUNCH_CHECK("amax_kernel");
  return OG_OK;
}

int og_linear_f16_fwd(const og_linear_args* a, const void* Wh16, const void* Wl16, const float* w_meta, const float* a_amax,
                      float* amax_out, float* scale_out, void* Yh, void* Yl, void* Yth, void* Ytl, int swap_halves, void* stream) {
  OG_CHECK_ARG(a && a->A && Wh16 && Wl16 && w_meta && a_amax, "linear_f16: null pointer");
  OG_CHECK_ARG(a->rows > 0 && a->nout > 0 && a->batch > 0 && a->k1 > 0 && a->k2 >= 0, "linear_f16: bad sizes");
  OG_CHECK_ARG(!a->rscale && !a->Yt, "linear_f16: rscale / fp32 transposed outputs are not part of the fp16 form");
  F16LinearArgs g;
  memset(&g, 0, sizeof(g));
  g.A = a->A; g.lda = a->lda; g.strideA = a->strideA; g.A2 = a->A2; g.lda2 = a->lda2; g.strideA2 = a->strideA2; g.k1 = a->k1; g.k2 = a->k2;
  if (a->strideW && a->strideW % a->ldw != 0) return fail(OG_EUNSUPPORTED, "linear_f16: strideW must be a multiple of ldw");
  g.b_rows_per_batch = a->strideW ? (int)(a->strideW / a->ldw) : 0;
  g.bias = a->bias; g.rows = a->rows; g.nout = a->nout; g.batch = a->batch; g.alpha = a->alpha; g.relu = a->relu;
  g.R = a->R; g.ldr = a->ldr; g.strideR = a->strideR;
  g.Y = a->Y; g.ldy = a->ldy; g.strideY = a->strideY;
  g.Yh = static_cast<__half*>(Yh); g.Yl = static_cast<__half*>(Yl);
  g.Yth = static_cast<__half*>(Yth); g.Ytl = static_cast<__half*>(Ytl); g.ldyt = a->ldyt; g.strideYt = a->strideYt;
  g.amax_in[0] = a_amax; g.w_meta = w_meta; g.amax_out = amax_out; g.scale_out = scale_out; g.swap_halves = swap_halves;
  const __half* bh = static_cast<const __half*>(Wh16); const __half* bl = static_cast<const __half*>(Wl16);
  if (!linear_f16_eligible(g, bh, bl, a->ldw))
    return fail(OG_EUNSUPPORTED, "linear_f16: needs K >= 64, 16-byte aligned rows, exactly one output kind (Y | Yh,Yl | Yth,Ytl)");
  const int64_t brows = a->strideW ? (int64_t)g.b_rows_per_batch * (a->batch - 1) + a->nout : a->nout;
  return linear_f16_launch(g, bh, bl, a->ldw, brows, (cudaStream_t)stream);
}

int og_attention_f16_fwd(const float* q, int64_t ldq, int64_t strideq, const float* q_amax, const void* khi, const void* klo,
                         int64_t ldk, const float* k_scale, const void* vthi, const void* vtlo, int64_t ldvt, const float* v_scale,
                         float* out, int64_t ldo, int64_t strideo, float* out_amax, int batch, int nq, int nk, int num_heads,
                         int head_dim, int swap_halves, void* stream) {
  OG_CHECK_ARG(q && q_amax && khi && klo && k_scale && vthi && vtlo && v_scale && out, "attention_f16: null pointer");
  OG_CHECK_ARG(batch > 0 && nq > 0 && nk > 0 && num_heads > 0, "attention_f16: bad sizes");
  if (!attention_f16_eligible(head_dim, ldq, ldk, ldvt, ldo))
    return fail(OG_EUNSUPPORTED, "attention_f16: head_dim 64 and 16-byte aligned rows required");
  TcAttnArgs a{q, ldq, strideq, out, ldo, strideo, batch, nq, nk, num_heads, num_heads * head_dim, (float)pow((double)head_dim, -0.5)};
  F16AttnScales sc{q_amax, k_scale, v_scale, out_amax, swap_halves};
  return attention_f16_dispatch(a, sc, static_cast<const __half*>(khi), static_cast<const __half*>(klo), ldk,
                                static_cast<const __half*>(vthi), static_cast<const __half*>(vtlo), ldvt, head_dim, (cudaStream_t)stream);
}

}  // extern "C"
