"""Debug: event timelines of CTA 0 (+ its peer) of the fp16 hi/lo kernels (libopenglue_b200_trace.so, -DOG_TRACE).
   python scripts/trace_f16.py attn | gemm
attention  MMA warps: 0 QK_i waits K | 1 issue QK_i | 2 PV_i waits | 3 P_i ready | 4 issue PV_i
           softmax  : 5 S_i observed | 8 row max exchanged | 9 exps + split done | 6 P_i handed over | 7 O_i folded
gemm       0 TMA issues block | 1 A landed | 2 A in TMEM | 3 B landed | 4 MMA issue | 5 chunk seen | 6 chunk drained | 8/9 epilogue start/end"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(here, 'openglue_b200', 'libopenglue_b200_trace.so'))
dev = 'cuda:0'
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
I, L = C.c_int, C.c_int64
which = sys.argv[1] if len(sys.argv) > 1 else 'attn'


def split16(x2d, bias=None):
    hi = torch.empty(x2d.shape, dtype=torch.float16, device=dev)
    lo, meta = torch.empty_like(hi), torch.zeros(4, device=dev)
    assert lib.og_weight_split_f16(p(x2d), p(bias), I(x2d.shape[0]), I(x2d.shape[1]), p(hi), p(lo), p(meta), st) == 0
    return hi, lo, meta


def amax(x):
    s = torch.zeros(1, device=dev)
    assert lib.og_amax(p(x), L(x.numel()), p(s), st) == 0
    return s


def read():
    buf = (C.c_longlong * (2 * 16 * 256))()
    assert lib.og_trace_read(buf) == 0
    return [[buf[e * 256 + i] for i in range(256)] for e in range(16)], [[buf[4096 + e * 256 + i] for i in range(256)] for e in range(16)]


if which == 'attn':
    nb, n, d, H = 32, 2048, 256, 4
    q = torch.randn(nb * n, d, device=dev); k = torch.randn(nb * n, d, device=dev); vt = torch.randn(nb * d, n, device=dev)
    kh, kl, km = split16(k); vh, vl, vm = split16(vt)
    o = torch.empty(nb * n, d, device=dev)
    qa = amax(q)
    for _ in range(3):
        rc = lib.og_attention_f16_fwd(p(q), L(d), L(n * d), p(qa), p(kh), p(kl), L(d), p(km), p(vh), p(vl), L(n), p(vm), p(o), L(d), L(n * d), None,
                                      I(nb), I(n), I(n), I(H), I(d // H), I(0), st)
        assert rc == 0, rc
    torch.cuda.synchronize()
    ev, ev1 = read()
    t0 = ev[1][0]
    print('leader  blk | waitK issQK | waitPV Pready issPV | S_seen maxdone expdone P_given O_folded || S->max max->exp exp->P  S->P  P->fold | dQK dPV  S_seen-issQK issPV-P_given')
    for i in range(4, 44):
        r = [ev[e][i] - t0 for e in range(10)]
        print(f'{i:3d} | {r[0]:6d} {r[1]:6d} | {r[2]:6d} {r[3]:6d} {r[4]:6d} | {r[5]:6d} {r[8]:6d} {r[9]:6d} {r[6]:6d} {r[7]:6d} || '
              f'{ev[8][i]-ev[5][i]:5d} {ev[9][i]-ev[8][i]:5d} {ev[6][i]-ev[9][i]:5d} {ev[6][i]-ev[5][i]:5d} {ev[7][i]-ev[6][i]:6d} | '
              f'{ev[1][i]-ev[1][i-1]:5d} {ev[4][i]-ev[4][i-1]:5d}  {ev[5][i]-ev[1][i]:6d} {ev[4][i]-ev[6][i]:6d}')
    print('peer CTA (own clock): S_seen->P_given, d(S_seen)')
    print(' '.join(f'{ev1[6][i]-ev1[5][i]}/{ev1[5][i]-ev1[5][i-1]}' for i in range(5, 30)))
else:
    rows, k1, nout = 65536, 512, 256          # the fc2 GEMM (residual through TMA)
    kind = sys.argv[2] if len(sys.argv) > 2 else 'fc2'
    if kind == 'q':
        k1, nout = 256, 256
    A = torch.randn(rows, k1, device=dev); W = torch.randn(nout, k1, device=dev) / 16; bias = torch.randn(nout, device=dev)
    Wh, Wl, meta = split16(W, bias)
    Y = torch.randn(rows, nout, device=dev)

    class Args(C.Structure):
        _fields_ = [('A', C.c_void_p), ('lda', L), ('strideA', L), ('A2', C.c_void_p), ('lda2', L), ('strideA2', L), ('k1', C.c_int32), ('k2', C.c_int32),
                    ('W', C.c_void_p), ('ldw', L), ('strideW', L), ('bias', C.c_void_p), ('rows', C.c_int32), ('nout', C.c_int32), ('batch', C.c_int32),
                    ('alpha', C.c_float), ('relu', C.c_int32), ('R', C.c_void_p), ('ldr', L), ('strideR', L), ('rscale', C.c_void_p),
                    ('Y', C.c_void_p), ('ldy', L), ('strideY', L), ('Yt', C.c_void_p), ('ldyt', L), ('strideYt', L)]
    a = Args()
    a.A, a.lda, a.k1, a.ldw, a.bias = A.data_ptr(), k1, k1, k1, bias.data_ptr()
    a.rows, a.nout, a.batch, a.alpha = rows, nout, 1, 1.0
    a.Y, a.ldy = Y.data_ptr(), nout
    if kind == 'fc2':
        a.R, a.ldr = Y.data_ptr(), nout
    am, ao = amax(A), torch.zeros(1, device=dev)
    for _ in range(3):
        rc = lib.og_linear_f16_fwd(C.byref(a), p(Wh), p(Wl), p(meta), p(am), p(ao), None, None, None, None, None, I(0), st)
        assert rc == 0, rc
    torch.cuda.synchronize()
    ev, ev1 = read()
    t0 = ev[0][0]
    nkb = k1 // 64
    print(f'GEMM {kind}: rows {rows} K {k1} nout {nout}; {nkb} K blocks per tile')
    print('kblk | tma  Aland  Atmem | Bland  mma | d(mma) | chunk_seen drained d(drained)')
    for i in range(2, 50):
        print(f'{i:3d} | {ev[0][i]-t0:6d} {ev[1][i]-t0:6d} {ev[2][i]-t0:6d} | {ev[3][i]-t0:6d} {ev[4][i]-t0:6d} | {ev[4][i]-ev[4][i-1]:5d} | {ev[5][i]-t0:6d} {ev[6][i]-t0:6d} {ev[6][i]-ev[6][i-1]:5d}')
    print('tile | epilogue start  end  (end-start)  d(start)')
    for t in range(1, 8):
        print(f'{t:3d} | {ev[8][t]-t0:7d} {ev[9][t]-t0:7d} {ev[9][t]-ev[8][t]:6d} {ev[8][t]-ev[8][t-1]:6d}')
