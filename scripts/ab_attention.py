"""A/B check of one build of the fp16 attention kernel (OG_LIB=<lib> python scripts/ab_attention.py): operator parity against a
torch float64 reference on the GPU at shapes that exercise several tiles per CTA pair, odd / even key-block counts, ragged rows
and keys, then CUDA-event timings of the self-layer and cross-layer launches of the headline configuration.  One JSON line."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openglue_b200 import _cabi  # noqa: E402

DEV = 'cuda:0'
lib = _cabi.lib()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def split16(x2d):
    hi = torch.empty(x2d.shape, dtype=torch.float16, device=DEV)
    lo, meta = torch.empty_like(hi), torch.zeros(4, device=DEV)
    _cabi.check(lib.og_weight_split_f16(p(x2d), None, x2d.shape[0], x2d.shape[1], p(hi), p(lo), p(meta), st()), 'split16')
    return hi, lo, meta


def setup(B, H, nq, nk, seed=2, dh=64):
    g = torch.Generator(device=DEV).manual_seed(seed)
    d = H * dh
    q = 3 * torch.randn(B, nq, d, generator=g, device=DEV)
    k = 3 * torch.randn(B, nk, d, generator=g, device=DEV)
    v = 3 * torch.randn(B, nk, d, generator=g, device=DEV)
    kh, kl, kmeta = split16(k.reshape(B * nk, d))
    ldvt = (nk + 7) // 8 * 8
    vt = torch.zeros(B * d, ldvt, device=DEV)
    vt[:, :nk] = v.transpose(1, 2).reshape(B * d, nk)
    vth, vtl, vmeta = split16(vt)
    qamax = torch.zeros(1, device=DEV)
    _cabi.check(lib.og_amax(p(q), q.numel(), p(qamax), st()), 'og_amax')
    out = torch.full((B, nq, d), float('nan'), device=DEV)
    oamax = torch.zeros(1, device=DEV)

    def run():
        _cabi.check(lib.og_attention_f16_fwd(p(q), d, nq * d, p(qamax), p(kh), p(kl), d, p(kmeta), p(vth), p(vtl), ldvt, p(vmeta),
                                             p(out), d, nq * d, p(oamax), B, nq, nk, H, dh, 0, st()), 'og_attention_f16_fwd')
    return q, k, v, out, oamax, run


def parity(B, H, nq, nk, pair=1):
    dh = 64
    q, k, v, out, oamax, run = setup(B, H, nq, nk)
    lib.og_set_tuning(-1, pair)
    try:
        run()
        torch.cuda.synchronize()
    finally:
        lib.og_set_tuning(-1, 1)
    hv = lambda t, n: t.double().reshape(B, n, H, dh).permute(0, 2, 1, 3)        # [B, H, n, dh]
    s = hv(q, nq) @ hv(k, nk).transpose(2, 3) / dh ** 0.5
    ref = (torch.softmax(s, dim=-1) @ hv(v, nk)).permute(0, 2, 1, 3).reshape(B, nq, H * dh)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    ok_amax = float(oamax) == float(out.abs().max())
    return err, ok_amax


def timing(B, H, nq, nk, reps=20):
    *_, run = setup(B, H, nq, nk, seed=5)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {'lib': os.environ.get('OG_LIB', 'default'), 'parity': {}, 'ok': True}
SHAPES = [(2, 4, 200, 333, 1), (1, 2, 65, 1, 1), (1, 4, 300, 130, 0),
          (10, 4, 1024, 1216, 1),      # 160 tiles on 74 CTA pairs: several tiles per pair, 19 key blocks (odd)
          (12, 4, 1000, 1100, 1),      # 192 tiles, 18 key blocks (even), ragged rows and keys
          (6, 4, 700, 64, 1),          # one key block per tile: only one team works
          (5, 4, 1024, 1216, 0)]       # single-CTA form, several tiles per CTA
try:
    for B, H, nq, nk, pair in SHAPES:
        err, ok_amax = parity(B, H, nq, nk, pair)
        res['parity'][f'{B}x{H}x{nq}x{nk}/pair{pair}'] = err
        if not (err <= 5e-6 and ok_amax):
            res['ok'] = False
    res['ms_self_32x4x2048x2048'] = timing(32, 4, 2048, 2048)
    res['ms_cross_16x4x2048x2048'] = timing(16, 4, 2048, 2048)
    res['tflops_self'] = 4.0 * 2048 * 2048 * 256 * 32 / (res['ms_self_32x4x2048x2048'] * 1e-3) / 1e12
except Exception as e:                                  # a trapped kernel poisons the context: report and stop
    res['ok'] = False
    res['error'] = repr(e)[:400]
print(json.dumps(res), flush=True)
