"""Sinkhorn backward pass: the reverse-recurrence oracle against what torch autograd gives for the
UNMODIFIED reference ``SuperGlue.get_matching_probs`` (tests/golden/sinkgrad_*.pt, minted by oracle/gen_golden_sinkhorn_grad.py), and the
CUDA path (og_sinkhorn_train_fwd / og_sinkhorn_bwd through openglue_b200.sinkhorn.matching_log_probs) against both.
Tolerance: gradients are sums of <= (N + M) T products of fp32 exponentials: 2e-4 of the largest gradient entry (the reference's own
fp32-vs-fp64 difference is 1e-6 of it), d dustbin 2e-4 relative."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sinkhorn_grad_oracle as SG                       # noqa: E402  (checker only)
from oracle.gen_golden_sinkhorn_grad import CASES, inputs           # noqa: E402  (input generator; no reference import at module level)
from oracle import loss_oracle as L                                 # noqa: E402

GOLDEN = list(CASES)


def _fx(name):
    fx = torch.load(os.path.join(ROOT, 'tests', 'golden', name + '.pt'), weights_only=False)
    S, dust, G, labels = inputs(fx['case'])
    return fx, S, dust, G, labels


def _upstream(fx, G, labels, scores_shape):
    return G.double() if G is not None else L.criterion_grad(labels, scores_shape)


@pytest.mark.parametrize('name', GOLDEN)
def test_oracle_matches_reference_autograd(name):
    fx, S, dust, G, labels = _fx(name)
    b, n, m, iters, reg = fx['case'][:5]
    Gd = _upstream(fx, G, labels, (b, n + 1, m + 1))
    scores = SG.forward_with_history(S.double(), dust.double(), iters, reg)[0]
    # (the reference builds norm / log_a / log_b in float32 even in a float64 run, superglue.py:98-101: 1e-7 on the scores)
    assert (scores - fx['scores_f64']).abs().max() <= 1e-6
    dS, dd = SG.backward(S.double(), dust.double(), iters, reg, Gd)
    scale = float(fx['dS_f64'].abs().max())
    assert (dS - fx['dS_f64']).abs().max() <= 1e-6 * scale
    assert abs(float(dd) - float(fx['ddustbin_f64'])) <= 1e-6 * max(abs(float(fx['ddustbin_f64'])), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize('name', GOLDEN)
def test_cuda_sinkhorn_backward_matches_reference_autograd(name):
    from openglue_b200.sinkhorn import matching_log_probs
    fx, S, dust, G, labels = _fx(name)
    b, n, m, iters, reg = fx['case'][:5]
    dev = 'cuda:0'
    s = S.to(dev).requires_grad_(True)
    d = dust.to(dev).requires_grad_(True)
    scores = matching_log_probs(s, d, iters, reg)
    assert (scores.detach().cpu().double() - fx['scores_f64']).abs().max() <= 2e-5
    Gd = _upstream(fx, G, labels, (b, n + 1, m + 1)).float().to(dev)
    scores.backward(Gd)
    scale = float(fx['dS_f64'].abs().max())
    err = float((s.grad.cpu().double() - fx['dS_f64']).abs().max())
    derr = abs(float(d.grad) - float(fx['ddustbin_f64']))
    print(f'\n[{name}] max|d dS| {err:.2e} of max|dS| {scale:.2e} ({err / scale:.1e} relative); d dustbin {float(d.grad):.6f} vs {float(fx["ddustbin_f64"]):.6f}')
    assert err <= 2e-4 * scale
    assert derr <= 2e-4 * max(abs(float(fx['ddustbin_f64'])), 1e-3)
    # deterministic
    s2 = S.to(dev).requires_grad_(True); d2 = dust.to(dev).requires_grad_(True)
    matching_log_probs(s2, d2, iters, reg).backward(Gd)
    assert torch.equal(s2.grad, s.grad) and torch.equal(d2.grad, d.grad)


@pytest.mark.gpu
def test_cuda_sinkhorn_backward_headline_shape():
    """N = M = 2048, T = 100 (one pair): against the reverse-recurrence oracle in float64 (the reference's autograd tape for this shape
    is ~3.4 GB per pair in fp64; the oracle itself is pinned to the reference on the golden cases above)."""
    from openglue_b200.sinkhorn import matching_log_probs
    g = torch.Generator().manual_seed(7)
    n = m = 2048
    S = 6.0 * torch.randn(1, n, m, generator=g)
    dust = torch.tensor(1.0)
    gt0 = torch.full((1, n), -1, dtype=torch.int64); gt1 = torch.full((1, m), -1, dtype=torch.int64)
    src, dst = torch.randperm(n, generator=g)[:1200], torch.randperm(m, generator=g)[:1200]
    gt0[0, src] = dst; gt1[0, dst] = src
    Gd = L.criterion_grad({'gt_matches0': gt0, 'gt_matches1': gt1}, (1, n + 1, m + 1))
    want_dS, want_dd = SG.backward(S.double(), dust.double(), 100, 1.0, Gd)
    dev = 'cuda:0'
    s = S.to(dev).requires_grad_(True); d = dust.to(dev).requires_grad_(True)
    matching_log_probs(s, d, 100, 1.0).backward(Gd.float().to(dev))
    scale = float(want_dS.abs().max())
    err = float((s.grad.cpu().double() - want_dS).abs().max())
    print(f'\n[headline shape] max|d dS| {err:.2e} of {scale:.2e}; d dustbin {float(d.grad):.6f} vs {float(want_dd):.6f}')
    assert err <= 2e-4 * scale and abs(float(d.grad) - float(want_dd)) <= 2e-4 * abs(float(want_dd))
