"""Debug: per-key-block event timeline of CTA 0 of the tcgen05 attention kernel (libopenglue_b200_trace.so, -DOG_TRACE).
 MMA warp: 0 wait K_i | 1 issue QK_i | 2 QK_{i+1} issued, wait P_i | 3 P_i ready | 4 issue PV_i
 softmax : 5 S_i observed | 6 P_i handed over | 7 O_i folded"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(here, 'openglue_b200', 'libopenglue_b200_trace.so'))
dev = 'cuda:0'
nb, n, d, H = 32, 2048, 256, 4
q = torch.randn(nb * n, d, device=dev); k = torch.randn(nb * n, d, device=dev); vt = torch.randn(nb * d, n, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
def split(t):
    hi, lo = torch.empty_like(t), torch.empty_like(t)
    lib.og_split_tf32(p(t), p(hi), p(lo), C.c_int64(t.numel()), st); return hi, lo
khi, klo = split(k); vthi, vtlo = split(vt)
o = torch.empty(nb * n, d, device=dev)
I, L = C.c_int, C.c_int64
for _ in range(3):
    rc = lib.og_attention_tc_fwd(p(q), L(d), L(n * d), p(khi), p(klo), L(d), p(vthi), p(vtlo), L(n), p(o), L(d), L(n * d),
                                 I(nb), I(n), I(n), I(H), I(d // H), st)
    assert rc == 0, rc
torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 16 * 256))()
assert lib.og_trace_read(buf) == 0
pair = os.environ.get('OG_ATTN_PAIR') != '0'
ev = [[buf[e * 256 + i] for i in range(256)] for e in range(8)]
t0 = ev[1][0]
print('leader CTA   blk  waitK  issQK | waitP  Pready issPV | S_seen P_given O_folded || softmax_lat  P_wait  dQK   dPV   S_seen-issQK  issPV-P_given')
for i in range(4, 40):
    r = [ev[e][i] - t0 for e in range(8)]
    print(f'{i:3d} {r[0]:6d} {r[1]:6d} | {r[2]:6d} {r[3]:6d} {r[4]:6d} | {r[5]:6d} {r[6]:6d} {r[7]:6d} || {ev[6][i]-ev[5][i]:6d} {ev[3][i]-ev[2][i]:6d} '
          f'{ev[1][i]-ev[1][i-1]:6d} {ev[4][i]-ev[4][i-1]:6d}   {ev[5][i]-ev[1][i]:6d} {ev[4][i]-ev[6][i]:6d}')
if pair:
    ev1 = [[buf[4096 + e * 256 + i] for i in range(256)] for e in range(8)]
    print('peer CTA (its own clock): S_seen P_given O_folded | softmax_lat  d(S_seen)')
    for i in range(4, 40):
        print(f'{i:3d} {ev1[5][i]-ev1[5][4]:7d} {ev1[6][i]-ev1[5][4]:7d} {ev1[7][i]-ev1[5][4]:7d} | {ev1[6][i]-ev1[5][i]:6d} {ev1[5][i]-ev1[5][i-1]:6d}')
