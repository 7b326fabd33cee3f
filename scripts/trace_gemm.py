"""Debug: per-k-block event timeline of CTA 0 of the persistent GEMM (needs libopenglue_b200_trace.so,
built with -DOG_TRACE).  Events (SM clock cycles, relative to the first):
  0 producer: stage free, about to issue TMA      1 converter: full[s] observed
  2 converter: A split stored in TMEM             3 MMA warp: full[s] observed   4 MMA warp: a_full observed (issue)
  5 epilogue: acc_full[chunk] observed            6 epilogue: chunk drained
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openglue_b200 import _cabi
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(here, 'openglue_b200', 'libopenglue_b200_trace.so'))
dev = 'cuda:0'
rows, k1, nout = 65536, 256, 768
A = torch.randn(rows, k1, device=dev); W = torch.randn(nout, k1, device=dev)
Whi, Wlo = torch.empty_like(W), torch.empty_like(W)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
lib.og_split_tf32(p(W), p(Whi), p(Wlo), C.c_int64(W.numel()), st)
Y = torch.empty(rows, nout, device=dev)
a = _cabi.OgLinearArgs()
a.A, a.lda, a.k1, a.k2, a.ldw = A.data_ptr(), k1, k1, 0, k1
a.rows, a.nout, a.batch, a.alpha = rows, nout, 1, 1.0
a.Y, a.ldy = Y.data_ptr(), nout
for _ in range(3):
    rc = lib.og_linear_tc_fwd(C.byref(a), p(Whi), p(Wlo), None, None, None, None, 2, st)
    assert rc == 0, rc
torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 16 * 256))()
assert lib.og_trace_read(buf) == 0
for cta in range(1 if os.environ.get('OG_GEMM_PAIR') == '0' else 2):
    ev = [[buf[cta * 4096 + e * 256 + i] for i in range(256)] for e in range(16)]
    t0 = ev[0][0]
    print(f'=== CTA {cta} (clock64 of its own SM)')
    print('kb   issue   a_land  a_empty_ok conv_done  mma_bfull mma_afull | d(issue) d(mma)  tma_lat  split+wait  st+arrive  conv_done->mma')
    for i in range(8, 72):
        r = [ev[e][i] - t0 for e in range(8)]
        print(f'{i:3d} {r[0]:8d} {r[1]:8d} {r[7]:9d} {r[2]:9d} {r[3]:9d} {r[4]:9d} | {ev[0][i]-ev[0][i-1]:6d} {ev[4][i]-ev[4][i-1]:6d}   '
              f'{ev[1][i]-ev[0][i]:6d} {ev[7][i]-ev[1][i]:8d} {ev[2][i]-ev[7][i]:8d} {ev[4][i]-ev[2][i]:10d}')
    print('chunks: acc_full seen, drained (delta to previous)')
    for g in range(4, 36):
        print(g, ev[5][g] - t0, ev[6][g] - ev[5][g], ev[5][g] - ev[5][g - 1])
    print('epilogue of warp 0 per tile (cycles after its start): barrier passed | cc0: enter staging, staged | cc1: same | stores issued')
    for t in range(1, 12):
        e = lambda k: ev[k][t] - ev[8][t]
        print(f'{t:3d} start={ev[8][t]-t0:8d} bar={e(9):5d} | cc0 {e(10):5d} {e(12):5d} | cc1 {e(11):5d} {e(13):5d} | end {e(14):5d}')
