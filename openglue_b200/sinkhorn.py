"""Differentiable dustbin-augmented log-domain Sinkhorn on the GPU: the forward AND the backward pass of the reference's
``SuperGlue.get_matching_probs`` + ``log_otp_solver`` (reference models/superglue/superglue.py:88-111,
models/superglue/optimal_transport.py:4-28) as hand-written kernels behind ``og_sinkhorn_train_fwd`` / ``og_sinkhorn_bwd``.

``matching_log_probs(S, dustbin_score, num_iters, reg)`` returns the [B, N+1, M+1] log-assignment and is a
``torch.autograd.Function``: its backward runs the T unrolled iterations in reverse from the scaling-vector history the
forward recorded (what torch autograd does for the reference in ``training_step``, models/matching_module.py:99-105, without
the T x (N+1)(M+1) tape).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi

__all__ = ['matching_log_probs']


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Sinkhorn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S: torch.Tensor, dustbin: torch.Tensor, num_iters: int, reg: float):
        dev = S.device
        if dev.type != 'cuda':
            raise RuntimeError('openglue_b200.matching_log_probs needs CUDA tensors (sm_100a); there is no CPU path')
        B, n, m = S.shape
        lds = (m + 3) // 4 * 4
        Sp = torch.zeros(B, n, lds, dtype=torch.float32, device=dev)       # 16-byte aligned rows (the kernels' layout)
        Sp[:, :, :m] = S.detach().float()
        dust = dustbin.detach().float().reshape(1).contiguous()
        lib = _cabi.lib()
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            scores = torch.empty(B, n + 1, m + 1, dtype=torch.float32, device=dev)
            hist = torch.empty(max(int(lib.og_sinkhorn_hist_floats(B, n, m, num_iters)), 1), dtype=torch.float32, device=dev)
            wsb = lib.og_sinkhorn_workspace_bytes(B, n, m)
            if wsb < 0:
                _cabi.check(int(wsb), 'og_sinkhorn_workspace_bytes')
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            _cabi.check(lib.og_sinkhorn_train_fwd(_p(Sp), lds, n * lds, _p(dust), B, n, m, int(num_iters), float(reg), _p(scores), _p(hist),
                                                  _p(ws), wsb, st), 'og_sinkhorn_train_fwd')
        ctx.save_for_backward(Sp, dust, hist)
        ctx.meta = (B, n, m, lds, int(num_iters), float(reg), dustbin.shape, S.dtype, dustbin.dtype)
        return scores

    @staticmethod
    def backward(ctx, G: torch.Tensor):
        Sp, dust, hist = ctx.saved_tensors
        B, n, m, lds, T, reg, dshape, sdt, ddt = ctx.meta
        dev = Sp.device
        G = G.detach().float().contiguous()
        lib = _cabi.lib()
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            dZ = torch.empty(B, n + 1, m + 1, dtype=torch.float32, device=dev)
            dd = torch.empty(1, dtype=torch.float32, device=dev)
            wsb = lib.og_sinkhorn_bwd_workspace_bytes(B, n, m, T)
            if wsb < 0:
                _cabi.check(int(wsb), 'og_sinkhorn_bwd_workspace_bytes')
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            _cabi.check(lib.og_sinkhorn_bwd(_p(Sp), lds, n * lds, _p(dust), B, n, m, T, reg, _p(hist), _p(G), _p(dZ), _p(dd), _p(ws), wsb, st),
                        'og_sinkhorn_bwd')
        return dZ[:, :n, :m].to(sdt), dd.reshape(dshape).to(ddt), None, None


def matching_log_probs(S: torch.Tensor, dustbin_score: torch.Tensor, num_iters: int, reg: float = 1.0) -> torch.Tensor:
    """reference SuperGlue.get_matching_probs(S) (superglue.py:88-111): S [B, N, M] -> log-assignment [B, N+1, M+1];
    differentiable with respect to ``S`` and ``dustbin_score``."""
    return _Sinkhorn.apply(S, dustbin_score, num_iters, reg)
