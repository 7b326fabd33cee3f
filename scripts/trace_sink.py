"""Debug: per-iteration phase timeline of CTAs 0 and 1 of the Sinkhorn launch (libopenglue_b200_trace.so, -DOG_TRACE).
   OG_LIB=openglue_b200/libopenglue_b200_trace.so python scripts/trace_sink.py B n m iters
events: 0 sweep start, 1 sweep end, 2 partials written (barrier arrive), 3 barrier passed, 4 v rebuilt"""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openglue_b200 import _cabi
B, n, m, iters = (int(x) for x in sys.argv[1:5])
dev = torch.device('cuda:0')
lib = _cabi.lib()
g = torch.Generator(device='cuda').manual_seed(3)
lds = (m + 3) // 4 * 4
S = torch.randn(B, n, lds, device=dev, generator=g) * 3
dust = torch.tensor([1.0], device=dev)
sc = torch.empty(B, n + 1, m + 1, device=dev)
wsb = lib.og_sinkhorn_workspace_bytes(B, n, m)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()
for _ in range(3):
    _cabi.check(lib.og_sinkhorn_fwd(p(S), lds, n * lds, p(dust), B, n, m, iters, 1.0, p(sc), p(ws), wsb, st), 'og_sinkhorn_fwd')
torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 16 * 256))()
lib.og_trace_read.argtypes = [C.c_void_p]
assert lib.og_trace_read(buf) == 0
names = ['sweep', 'reduce+write partials', 'barrier wait', 'rebuild v', 'to next sweep']
for cta in range(2):
    ev = lambda e, i: buf[cta * 4096 + e * 256 + i]
    nit = min(iters, 256)
    print('CTA %d: cycles per iteration (mean over iterations 10..%d)' % (cta, nit - 2))
    acc = [0.0] * 5; cnt = 0
    for i in range(10, nit - 1):
        d = [ev(1, i) - ev(0, i), ev(2, i) - ev(1, i), ev(3, i) - ev(2, i), ev(4, i) - ev(3, i), ev(0, i + 1) - ev(4, i)]
        for k in range(5): acc[k] += d[k]
        cnt += 1
    tot = sum(acc) / cnt
    for k in range(5): print('   %-24s %9.0f  (%4.1f %%)' % (names[k], acc[k] / cnt, 100 * acc[k] / cnt / tot))
    print('   %-24s %9.0f' % ('iteration', tot))
    for i in (20, 21, 22):
        print('   it %d:' % i, [ev(e, i) - ev(0, 20) for e in range(5)])

# row timeline of warp 0 (iteration 20): events 5 loop top, 6 row in registers, 7 after warp max, 8 after warp sum, 9 after the segment exchange, 10 row done
rn = ['wait + load row', 'add v, max (shuffles)', 'exp, sum (shuffles)', 'segment exchange (bar)', 'divide, column fma', 'loop back']
for cta in range(2):
    ev = lambda e, i: buf[cta * 4096 + e * 256 + i]
    rows = [i for i in range(1, 255) if ev(10, i) > 0 and ev(5, i + 1) > 0]
    if not rows: continue
    acc = [0.0] * 6
    for i in rows:
        d = [ev(6, i) - ev(5, i), ev(7, i) - ev(6, i), ev(8, i) - ev(7, i), ev(9, i) - ev(8, i), ev(10, i) - ev(9, i), ev(5, i + 1) - ev(10, i)]
        for k in range(6): acc[k] += d[k]
    tot = sum(acc) / len(rows)
    print('CTA %d warp 0, iteration 20: cycles per row over %d rows' % (cta, len(rows)))
    for k in range(6): print('   %-26s %7.0f  (%4.1f %%)' % (rn[k], acc[k] / len(rows), 100 * acc[k] / len(rows) / tot))
    print('   %-26s %7.0f' % ('row', tot))
    print('   first rows:', [[ev(e, i) - ev(5, 1) for e in range(5, 11)] for i in (1, 2, 3, 4)])
