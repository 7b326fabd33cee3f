// extern "C" entry points of libopenglue_b200.so (see include/openglue_b200.h) and the
// host-side schedule of the whole matching-core forward pass.
#include "common.cuh"
#include "linear_simt.cuh"
#include "linear_tc.cuh"
#include "linear_tc2.cuh"
#include "attention_simt.cuh"
#include "attention_tc.cuh"
#include "sinkhorn.cuh"
#include "match.cuh"
#include "gt_matches.cuh"
#include "criterion.cuh"
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
  w->ldn = align_up(n, 4); w->ldm = align_up(m, 4);
  if (c->precision == OG_PREC_TF32X3) {
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
  const int64_t brows = a.strideW ? (int64_t)t.b_rows_per_batch * a.batch : a.nout;
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

static int attention_dispatch(const AttnArgs& a, int head_dim, int precision, cudaStream_t s) {
  (void)precision;
  return attention_simt_launch(a, head_dim, s);
}

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
  AttnArgs a{q, ldq, strideq, k, ldk, stridek, v, ldv, stridev, out, ldo, strideo, batch, nq, nk, num_heads,
             (float)pow((double)head_dim, -0.5)};
  return attention_dispatch(a, head_dim, precision, (cudaStream_t)stream);
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

int64_t og_criterion_workspace_bytes(int batch) { return batch > 0 ? criterion_workspace_bytes(batch) : -1; }

int og_criterion_fwd(const float* scores, const int64_t* gt_matches0, const int64_t* gt_matches1, int batch, int n, int m,
                     float* loss, float* dscores, float grad_scale, void* workspace, int64_t workspace_bytes, void* stream) {
  OG_CHECK_ARG(scores && gt_matches0 && gt_matches1 && loss && workspace, "criterion: null pointer");
  OG_CHECK_ARG(batch > 0 && n > 0 && m > 0, "criterion: bad sizes");
  return criterion_launch(scores, gt_matches0, gt_matches1, batch, n, m, loss, dscores, grad_scale, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int og_superglue_forward(const og_config* cfg, const float* Wp, const float* Whi, const float* Wlo, int B, int n, int m,
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
  OG_CHECK_ARG(cfg->precision == OG_PREC_FP32 || (Whi && Wlo), "forward: OG_PREC_TF32X3 needs packed_hi / packed_lo");
  const bool tcp = (cfg->precision == OG_PREC_TF32X3) && cfg->descriptor_dim >= 32;   // K >= 32 for the tcgen05 tiles
  auto WH = [&](int64_t off) { return tcp ? Whi + off : nullptr; };
  auto WL = [&](int64_t off) { return tcp ? Wlo + off : nullptr; };
  OG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "forward: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream_;
  const int prec = cfg->precision;
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
    return attention_dispatch(a, dh, prec, st);
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
  for (int l = 0; l < cfg->num_layers; ++l) {
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

}  // extern "C"
