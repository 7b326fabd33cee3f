"""GPU probe for the tcgen05 3xTF32 GEMM (og_linear_tc_fwd): each case runs in its own process
with a timeout, so a trap or hang in one variant does not hide the others.
    python scripts/tc_probe.py            # prints one line per case
"""
import json
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r'''
import sys, json, ctypes as C, torch
sys.path.insert(0, %(root)r)
from openglue_b200 import _cabi
mode, rows, k1, k2, nout, batch, bstride, relu, resid, split, transposed, reps = %(args)r
dev = 'cuda:0'
g = torch.Generator().manual_seed(0)
K = k1 + k2
A = torch.randn(batch, rows, k1, generator=g) * 3
A2 = torch.randn(batch, rows, k2, generator=g) if k2 else None
nb = batch if bstride else 1
W = torch.randn(nb, nout, K, generator=g)
bias = torch.randn(nout, generator=g)
R = torch.randn(batch, rows, nout, generator=g) if resid else None
X = torch.cat([A, A2], -1) if k2 else A
ref = 0.5 * (X.double() @ W.double().transpose(1, 2)) + bias.double()
if relu: ref = ref.relu()
if resid: ref = ref + R.double()
lib = _cabi.lib()
dA, dW, db = A.to(dev), W.to(dev), bias.to(dev)
dA2 = A2.to(dev) if k2 else None
dR = R.to(dev) if resid else None
Whi, Wlo = torch.empty_like(dW), torch.empty_like(dW)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
_cabi.check(lib.og_split_tf32(p(dW), p(Whi), p(Wlo), dW.numel(), st), 'split')
assert (Whi.double() + Wlo.double() - dW.double()).abs().max() <= dW.abs().max() * 2**-21
Y = torch.full((batch, rows, nout), float('nan'), device=dev)
Yhi = torch.zeros_like(Y) if split else None; Ylo = torch.zeros_like(Y) if split else None
Yt = torch.full((batch, nout, rows), float('nan'), device=dev) if transposed else None
Ythi = torch.zeros_like(Yt) if (split and transposed) else None; Ytlo = torch.zeros_like(Yt) if (split and transposed) else None
a = _cabi.OgLinearArgs()
a.A, a.lda, a.strideA = dA.data_ptr(), k1, rows * k1
if k2: a.A2, a.lda2, a.strideA2 = dA2.data_ptr(), k2, rows * k2
a.k1, a.k2 = k1, k2
a.W, a.ldw, a.strideW = 0, K, (nout * K if bstride else 0)
a.bias = db.data_ptr()
a.rows, a.nout, a.batch, a.alpha, a.relu = rows, nout, batch, 0.5, int(relu)
if resid: a.R, a.ldr, a.strideR = dR.data_ptr(), nout, rows * nout
a.Y, a.ldy, a.strideY = Y.data_ptr(), nout, rows * nout
if transposed: a.Yt, a.ldyt, a.strideYt = Yt.data_ptr(), rows, nout * rows
def run():
    _cabi.check(lib.og_linear_tc_fwd(C.byref(a), p(Whi), p(Wlo), p(Yhi), p(Ylo), p(Ythi), p(Ytlo), mode, st), 'og_linear_tc_fwd')
run(); torch.cuda.synchronize()
scale = float(ref.abs().max())
out = {'err_rel': float((Y.cpu().double() - ref).abs().max()) / scale, 'nan': int(torch.isnan(Y).sum())}
if split: out['split_err'] = float((Yhi.double() + Ylo.double() - Y.double()).abs().max()) / scale
if transposed: out['t_ok'] = bool(torch.equal(Yt.transpose(1, 2), Y))
if split and transposed: out['tsplit_ok'] = bool(torch.equal(Ythi.transpose(1, 2), Yhi) and torch.equal(Ytlo.transpose(1, 2), Ylo))
if reps:
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out['ms'] = ms; out['tflops'] = 2.0 * batch * rows * K * nout / ms / 1e9
print('RESULT ' + json.dumps(out))
'''

# mode(0=TS,1=SS), rows, k1, k2, nout, batch, bstride(B per batch), relu, resid, split, transposed, timing reps
CASES = [
    ('ss_one_kblock',   (1, 128, 32, 0, 128, 1, 0, 0, 0, 0, 0, 0)),
    ('ts_one_kblock',   (0, 128, 32, 0, 128, 1, 0, 0, 0, 0, 0, 0)),
    ('ss_k256',         (1, 128, 256, 0, 128, 1, 0, 0, 0, 0, 0, 0)),
    ('ts_k256',         (0, 128, 256, 0, 128, 1, 0, 0, 0, 0, 0, 0)),
    ('ts_tails_concat', (0, 1000, 256, 256, 392, 1, 0, 1, 0, 0, 0, 0)),
    ('ts_resid_split_t', (0, 300, 512, 0, 256, 2, 0, 0, 1, 1, 1, 0)),
    ('ts_batchedB',     (0, 257, 64, 0, 200, 3, 1, 0, 0, 0, 0, 0)),
    ('ss_tails_concat', (1, 1000, 256, 256, 392, 1, 0, 1, 0, 0, 0, 0)),
    ('v2_k256',         (2, 128, 256, 0, 128, 1, 0, 0, 0, 0, 0, 0)),
    ('v2_tails_concat', (2, 1000, 256, 256, 392, 1, 0, 1, 0, 0, 0, 0)),
    ('v2_resid_split_t', (2, 300, 512, 0, 256, 2, 0, 0, 1, 1, 1, 0)),
    ('v2_batchedB',     (2, 257, 64, 0, 200, 3, 1, 0, 0, 0, 0, 0)),
    ('v2_qkv_shape_time', (2, 65536, 256, 0, 768, 1, 0, 0, 0, 0, 0, 10)),
    ('v2_fc1_shape_time', (2, 65536, 256, 256, 512, 1, 0, 1, 0, 0, 0, 10)),
    ('v2_fc2_shape_time', (2, 65536, 512, 0, 256, 1, 0, 0, 1, 0, 0, 10)),
    ('v2_score_shape_time', (2, 2048, 256, 0, 2048, 16, 1, 0, 0, 0, 0, 10)),
    ('ts_qkv_shape_time', (0, 65536, 256, 0, 768, 1, 0, 0, 0, 0, 0, 10)),
    ('ts_fc1_shape_time', (0, 65536, 256, 256, 512, 1, 0, 1, 0, 0, 0, 10)),
    ('ss_qkv_shape_time', (1, 65536, 256, 0, 768, 1, 0, 0, 0, 0, 0, 10)),
]

if __name__ == '__main__':
    only = sys.argv[1:]
    for name, args in CASES:
        if only and name not in only:
            continue
        code = CASE % {'root': ROOT, 'args': args}
        try:
            r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=180)
            res = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
            if res:
                print(f'{name:22s} {res[0][7:]}', flush=True)
            else:
                tail = (r.stdout + r.stderr).strip().splitlines()[-6:]
                print(f'{name:22s} FAILED rc={r.returncode}: ' + ' | '.join(tail), flush=True)
        except subprocess.TimeoutExpired:
            print(f'{name:22s} TIMEOUT', flush=True)
