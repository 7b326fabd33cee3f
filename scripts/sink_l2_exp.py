"""Sinkhorn operator timing for one OG_SINK_L2_MB setting (read once per process by the library): prints ms and algorithmic GB/s.
usage: OG_SINK_L2_MB=80 python scripts/sink_l2_exp.py B n m iters"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openglue_b200 import _cabi

B, n, m, iters = (int(x) for x in sys.argv[1:5])
dev = torch.device('cuda:0')
torch.cuda.init(); torch.zeros(1, device=dev)
if os.environ.get('OG_PERSIST_MB'):
    import ctypes as C
    rt = C.CDLL('libcudart.so')
    val = C.c_int()
    rt.cudaDeviceGetAttribute(C.byref(val), 108, 0)        # cudaDevAttrMaxPersistingL2CacheSize
    want = min(int(os.environ['OG_PERSIST_MB']) << 20, val.value)
    rc = rt.cudaDeviceSetLimit(6, C.c_size_t(want))       # cudaLimitPersistingL2CacheSize
    got = C.c_size_t()
    rt.cudaDeviceGetLimit(C.byref(got), 6)
    print('persisting L2: max %d MB, set rc=%d, limit now %d MB' % (val.value >> 20, rc, got.value >> 20), flush=True)
lib = _cabi.lib()
g = torch.Generator(device='cuda').manual_seed(3)
lds = (m + 3) // 4 * 4
S = torch.randn(B, n, lds, device=dev, generator=g) * 3
dust = torch.tensor([1.0], device=dev)
sc = torch.empty(B, n + 1, m + 1, device=dev)
wsb = lib.og_sinkhorn_workspace_bytes(B, n, m)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()
def run():
    _cabi.check(lib.og_sinkhorn_fwd(p(S), lds, n * lds, p(dust), B, n, m, iters, 1.0, p(sc), p(ws), wsb, st), 'og_sinkhorn_fwd')
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
byts = B * (n * m * 4.0 * (iters + 1) + (n + 1) * (m + 1) * 4.0)
def torch_reference(Sb, iters):          # superglue.py:88-111 + optimal_transport.py:4-28 in fp64 (checker for the experiment configurations)
    import math
    Bq, nn, mm = Sb.shape
    Z = torch.full((Bq, nn + 1, mm + 1), 1.0, dtype=torch.float64, device=dev)
    Z[:, :nn, :mm] = Sb.double()
    norm = -math.log(nn + mm)
    la = torch.full((nn + 1,), norm, dtype=torch.float64, device=dev); la[-1] += math.log(mm)
    lb = torch.full((mm + 1,), norm, dtype=torch.float64, device=dev); lb[-1] += math.log(nn)
    u = torch.zeros(Bq, nn + 1, dtype=torch.float64, device=dev); v = torch.zeros(Bq, mm + 1, dtype=torch.float64, device=dev)
    for _ in range(iters):
        u = la - torch.logsumexp(Z + v[:, None, :], 2)
        v = lb - torch.logsumexp(Z + u[:, :, None], 1)
    return Z + u[:, :, None] + v[:, None, :] - norm
err = float((sc[:1].double() - torch_reference(S[:1, :, :m], iters)).abs().max())
print('cfg=%s max |scores - fp64 reference| (pair 0) %.2e' % (os.environ.get('OG_SINK_CFG', 'default'), err), flush=True)
print('L2_MB=%s B=%d n=%d m=%d iters=%d  ms=%.3f  GB/s=%.0f  checksum=%.6f' % (os.environ.get('OG_SINK_L2_MB', 'default'), B, n, m, iters, ms,
      byts / ms / 1e6, sc.double().sum().item()), flush=True)
