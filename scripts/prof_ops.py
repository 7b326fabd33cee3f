"""Run the dominant kernels once each at the headline (C3) shapes - the target of
`ncu --set full -k regex:...` captures.   python scripts/prof_ops.py [attn|xattn|qkv|linear|sinkhorn|all] [reps]
(xattn = one cross-attention launch: 16 sequences, N x M; qkv = the three projection GEMM variants of one layer:
Q -> fp32 Y, K -> split row-major, V -> split transposed)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openglue_b200 import _cabi

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = 'cuda:0'
lib = _cabi.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
B, n, d, H = 16, 2048, 256, 4
torch.manual_seed(0)


def split(t):
    hi, lo = torch.empty_like(t), torch.empty_like(t)
    _cabi.check(lib.og_split_tf32(p(t), p(hi), p(lo), t.numel(), st), 'split')
    return hi, lo


if which in ('attn', 'all'):
    nb = 2 * B
    q = torch.randn(nb * n, d, device=dev)
    k = torch.randn(nb * n, d, device=dev)
    vt = torch.randn(nb * d, n, device=dev)
    khi, klo = split(k)
    vthi, vtlo = split(vt)
    o = torch.empty(nb * n, d, device=dev)
    for _ in range(reps):
        _cabi.check(lib.og_attention_tc_fwd(p(q), d, n * d, p(khi), p(klo), d, p(vthi), p(vtlo), n, p(o), d, n * d,
                                            nb, n, n, H, d // H, st), 'attn')
if which in ('xattn', 'evidence'):
    q = torch.randn(B * n, d, device=dev)
    k = torch.randn(B * n, d, device=dev)
    vt = torch.randn(B * d, n, device=dev)
    khi, klo = split(k)
    vthi, vtlo = split(vt)
    o = torch.empty(B * n, d, device=dev)
    for _ in range(reps):
        _cabi.check(lib.og_attention_tc_fwd(p(q), d, n * d, p(khi), p(klo), d, p(vthi), p(vtlo), n, p(o), d, n * d,
                                            B, n, n, H, d // H, st), 'xattn')
if which in ('qkv', 'evidence'):
    rows = 2 * B * n
    A = torch.randn(rows, d, device=dev)
    W = torch.randn(d, d, device=dev) / 16
    Whi, Wlo = split(W)
    bias = torch.randn(d, device=dev)
    Y = torch.empty(rows, d, device=dev)
    Yhi, Ylo = torch.empty_like(Y), torch.empty_like(Y)
    Yt_hi, Yt_lo = torch.empty(2 * B, d, n, device=dev), torch.empty(2 * B, d, n, device=dev)
    for kind in ('q', 'k', 'v'):
        a = _cabi.OgLinearArgs()
        a.k1, a.k2, a.ldw, a.strideW = d, 0, d, 0
        a.bias = bias.data_ptr()
        a.nout, a.alpha, a.relu = d, 1.0, 0
        if kind == 'v':                                  # per sequence: transposed split output [B, d, n]
            a.A, a.lda, a.strideA = A.data_ptr(), d, n * d
            a.rows, a.batch, a.ldyt, a.strideYt, a.ldy = n, 2 * B, n, d * n, d
            outs = (None, None, p(Yt_hi), p(Yt_lo))
        else:
            a.A, a.lda, a.strideA = A.data_ptr(), d, 0
            a.rows, a.batch, a.ldy, a.strideY = rows, 1, d, 0
            if kind == 'q':
                a.Y = Y.data_ptr()
            outs = (None, None, None, None) if kind == 'q' else (p(Yhi), p(Ylo), None, None)
        for _ in range(reps):
            _cabi.check(lib.og_linear_tc_fwd(C.byref(a), p(Whi), p(Wlo), *outs, 2, st), 'qkv ' + kind)
if which in ('f16', 'evidence16'):
    # the fp16 hi/lo kernels at the headline shapes: self attention, cross attention, Q / K / V projections, fc1, fc2
    def split16(x2d, bias=None):
        hi = torch.empty(x2d.shape, dtype=torch.float16, device=dev)
        lo, meta = torch.empty_like(hi), torch.zeros(4, device=dev)
        _cabi.check(lib.og_weight_split_f16(p(x2d), p(bias), x2d.shape[0], x2d.shape[1], p(hi), p(lo), p(meta), st), 'split16')
        return hi, lo, meta

    def amax(x):
        s_ = torch.zeros(1, device=dev)
        _cabi.check(lib.og_amax(p(x), x.numel(), p(s_), st), 'amax')
        return s_
    for nb in (2 * B, B):                                   # self layer (32 sequences), cross layer (16)
        q = torch.randn(nb * n, d, device=dev); k = torch.randn(nb * n, d, device=dev); vt = torch.randn(nb * d, n, device=dev)
        kh, kl, km = split16(k); vh, vl, vm = split16(vt)
        o = torch.empty(nb * n, d, device=dev)
        qa = amax(q)
        for _ in range(reps):
            _cabi.check(lib.og_attention_f16_fwd(p(q), d, n * d, p(qa), p(kh), p(kl), d, p(km), p(vh), p(vl), n, p(vm), p(o), d, n * d, None,
                                                 nb, n, n, H, d // H, 0, st), 'attention_f16')
    rows = 2 * B * n
    for (kind, k1, k2, nout, relu, resid) in (('q', 256, 0, 256, 0, 0), ('k', 256, 0, 256, 0, 0), ('v', 256, 0, 256, 0, 0),
                                              ('fc1', 256, 256, 512, 1, 0), ('fc2', 512, 0, 256, 0, 1)):
        A = torch.randn(rows, k1, device=dev)
        A2 = torch.randn(rows, k2, device=dev) if k2 else None
        W = torch.randn(nout, k1 + k2, device=dev) / 16
        bias = torch.randn(nout, device=dev)
        Wh, Wl, meta = split16(W, bias)
        am = amax(torch.cat([A.flatten(), A2.flatten()]) if k2 else A)
        Y = torch.randn(rows, nout, device=dev)
        a = _cabi.OgLinearArgs()
        a.k1, a.k2, a.ldw, a.strideW = k1, k2, k1 + k2, 0
        a.bias = bias.data_ptr()
        a.nout, a.alpha, a.relu = nout, 1.0, relu
        a.A, a.lda = A.data_ptr(), k1
        if k2:
            a.A2, a.lda2 = A2.data_ptr(), k2
        ao, so = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        outs = [None, None, None, None]
        if kind == 'v':
            a.strideA, a.rows, a.batch, a.ldyt, a.strideYt = n * k1, n, 2 * B, n, nout * n
            yt_h = torch.empty(2 * B, nout, n, dtype=torch.float16, device=dev); yt_l = torch.empty_like(yt_h)
            outs[2], outs[3] = p(yt_h), p(yt_l)
        else:
            a.rows, a.batch, a.ldy = rows, 1, nout
            if kind == 'k':
                yh = torch.empty(rows, nout, dtype=torch.float16, device=dev); yl = torch.empty_like(yh)
                outs[0], outs[1] = p(yh), p(yl)
            else:
                a.Y = Y.data_ptr()
                if resid:
                    a.R, a.ldr = Y.data_ptr(), nout
        for _ in range(reps):
            _cabi.check(lib.og_linear_f16_fwd(C.byref(a), p(Wh), p(Wl), p(meta), p(am), p(ao) if a.Y else None,
                                              None if a.Y else p(so), *outs, 0, st), 'linear_f16 ' + kind)
if which in ('linear', 'all'):
    rows = 2 * B * n
    for (k1, k2, nout, relu, resid) in ((256, 0, 256, 0, 0), (256, 256, 512, 1, 0), (512, 0, 256, 0, 1)):
        A = torch.randn(rows, k1, device=dev)
        A2 = torch.randn(rows, k2, device=dev) if k2 else None
        W = torch.randn(nout, k1 + k2, device=dev)
        Whi, Wlo = split(W)
        bias = torch.randn(nout, device=dev)
        Y = torch.randn(rows, nout, device=dev)
        a = _cabi.OgLinearArgs()
        a.A, a.lda, a.strideA = A.data_ptr(), k1, 0
        if k2:
            a.A2, a.lda2, a.strideA2 = A2.data_ptr(), k2, 0
        a.k1, a.k2, a.ldw, a.strideW = k1, k2, k1 + k2, 0
        a.bias = bias.data_ptr()
        a.rows, a.nout, a.batch, a.alpha, a.relu = rows, nout, 1, 1.0, relu
        if resid:
            a.R, a.ldr, a.strideR = Y.data_ptr(), nout, 0
        a.Y, a.ldy, a.strideY = Y.data_ptr(), nout, 0
        for _ in range(reps):
            _cabi.check(lib.og_linear_tc_fwd(C.byref(a), p(Whi), p(Wlo), None, None, None, None, 2, st), 'linear')
if which in ('sinkhorn', 'all'):
    S = torch.randn(B, n, n, device=dev) * 4
    scores = torch.empty(B, n + 1, n + 1, device=dev)
    dust = torch.ones(1, device=dev)
    wsb = lib.og_sinkhorn_workspace_bytes(B, n, n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    for _ in range(reps):
        _cabi.check(lib.og_sinkhorn_fwd(p(S), n, n * n, p(dust), B, n, n, 100, 1.0, p(scores), p(ws), wsb, st), 'sinkhorn')
torch.cuda.synchronize()
print('done')
