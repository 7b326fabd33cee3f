"""CPU oracle for the Sinkhorn backward pass  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates in plain torch (any dtype; float64 in the tests) the reverse recurrences csrc/sinkhorn_bwd.cuh implements: the gradient of
the reference's ``SuperGlue.get_matching_probs`` / ``log_otp_solver`` (superglue.py:88-111, optimal_transport.py:4-28) through all
T unrolled iterations, from the scaling-vector history alone.  Pinned by tests/golden/sinkgrad_*.pt, which hold what torch autograd
gives for the UNMODIFIED reference (oracle/gen_golden_sinkhorn_grad.py); tests/test_sinkhorn_grad.py checks every fixture.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

Tensor = torch.Tensor


def forward_with_history(S: Tensor, dustbin: Tensor, iters: int, reg: float):
    """superglue.py:88-111 + optimal_transport.py:20-28, keeping every u_t / v_t."""
    B, n, m = S.shape
    Z = torch.empty(B, n + 1, m + 1, dtype=S.dtype)
    Z[:, :n, :m] = S
    Z[:, n, :] = dustbin
    Z[:, :, m] = dustbin
    norm = -math.log(n + m)
    log_a = torch.full((n + 1,), norm, dtype=S.dtype); log_a[-1] += math.log(m)
    log_b = torch.full((m + 1,), norm, dtype=S.dtype); log_b[-1] += math.log(n)
    Z = Z / reg
    u = torch.zeros(B, n + 1, dtype=S.dtype)
    v = torch.zeros(B, m + 1, dtype=S.dtype)
    us, vs = [], [v]
    for _ in range(iters):
        u = log_a - torch.logsumexp(Z + v[:, None, :], dim=2)
        v = log_b - torch.logsumexp(Z + u[:, :, None], dim=1)
        us.append(u); vs.append(v)
    scores = Z + u[:, :, None] + v[:, None, :] - norm
    return scores, Z, us, vs, log_a, log_b


def backward(S: Tensor, dustbin: Tensor, iters: int, reg: float, G: Tensor) -> Tuple[Tensor, Tensor]:
    """-> (d loss / d S [B, n, m], d loss / d dustbin) for G = d loss / d scores [B, n+1, m+1]."""
    B, n, m = S.shape
    _, Z, us, vs, log_a, log_b = forward_with_history(S, dustbin, iters, reg)
    ubar = G.sum(2)                                          # scores = Z + u_T + v_T - norm
    vbar = G.sum(1)
    Zbar = G.clone()
    for t in range(iters, 0, -1):
        u_t, v_t, v_tm1 = us[t - 1], vs[t], vs[t - 1]
        cvec = v_t - log_b                                   # v_t = log_b - LSE_i(Z + u_t)
        E = torch.exp(Z + u_t[:, :, None] + cvec[:, None, :])            # P2: column-normalised
        ub = (ubar if t == iters else torch.zeros_like(ubar)) - (E * vbar[:, None, :]).sum(2)
        coef = ub * torch.exp(-log_a)                        # u_t = log_a - LSE_j(Z + v_{t-1}):  P1 = E wq_j / a_i
        wq = torch.exp(v_tm1 - cvec)
        Zbar = Zbar - E * (vbar[:, None, :] + coef[:, :, None] * wq[:, None, :])
        vbar = -wq * (E * coef[:, :, None]).sum(1)
    dZ = Zbar / reg
    ddust = dZ[:, n, :].sum() + dZ[:, :n, m].sum()
    return dZ[:, :n, :m], ddust
