"""Mint golden vectors for the Sinkhorn backward pass: the UNMODIFIED reference ``SuperGlue.get_matching_probs``
(ucuapps/OpenGlue @ /root/reference, models/superglue/superglue.py:88-111 -> optimal_transport.py:4-28) under torch
autograd, for a dense upstream gradient and for the sparse one the reference's criterion produces.

TEST INFRASTRUCTURE.  Runs only in the build container; outputs: tests/golden/sinkgrad_*.pt (tests/test_sinkhorn_grad.py).

    python oracle/gen_golden_sinkhorn_grad.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('OPENGLUE_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

CASES = {
    # name: (batch, n, m, iters, reg, score scale, seed, upstream)
    'sinkgrad_small':  (2, 37, 52, 20, 1.0, 3.0, 1, 'dense'),
    'sinkgrad_reg':    (1, 64, 40, 7, 0.7, 2.0, 2, 'dense'),
    'sinkgrad_loss':   (2, 150, 131, 50, 1.0, 6.0, 3, 'criterion'),
    'sinkgrad_wide':   (1, 20, 1100, 10, 1.0, 2.0, 4, 'dense'),
}


def inputs(case):
    b, n, m, iters, reg, scale, seed, upstream = case
    g = torch.Generator().manual_seed(seed)
    S = scale * torch.randn(b, n, m, generator=g)
    dust = torch.tensor(0.5 + 0.1 * seed)
    if upstream == 'dense':
        G = torch.randn(b, n + 1, m + 1, generator=g)
        labels = None
    else:                               # the gradient of the reference's criterion for planted labels (utils/losses.py:7-53)
        from oracle.gen_golden_loss import synthetic_labels
        gt0, gt1, _ = synthetic_labels(b, n, m, seed, 0.6, 0.05, False)
        labels = {'gt_matches0': gt0, 'gt_matches1': gt1}
        G = None
    return S, dust, G, labels


def main():
    from models.superglue.superglue import SuperGlue              # the reference, unmodified
    from utils.losses import criterion
    from openglue_b200.synthetic import default_config
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, case in CASES.items():
        b, n, m, iters, reg, scale, seed, upstream = case
        S, dust, G, labels = inputs(case)
        cfg = default_config(descriptor_dim=32, num_stages=1, num_iters=iters, reg=reg)
        fx = {'case': case, 'reference': 'SuperGlue.get_matching_probs @ /root/reference under autograd, torch ' + torch.__version__}
        for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            model = SuperGlue(cfg).to(dtype)
            with torch.no_grad():
                model.dustbin_score.copy_(dust.to(dtype))
            s = S.to(dtype).clone().requires_grad_(True)
            scores = model.get_matching_probs(s)
            if upstream == 'dense':
                (scores * G.to(dtype)).sum().backward()
            else:
                y_pred = {'context_descriptors0': torch.zeros(b, 4, n, dtype=dtype), 'context_descriptors1': torch.zeros(b, 4, m, dtype=dtype),
                          'scores': scores}
                criterion(labels, y_pred, margin=None)['loss'].backward()
            fx[f'scores_{tag}'] = scores.detach().clone()
            fx[f'dS_{tag}'] = s.grad.detach().clone()
            fx[f'ddustbin_{tag}'] = model.dustbin_score.grad.detach().clone()
        torch.save(fx, os.path.join(out_dir, name + '.pt'))
        print(f'{name}: max|dS| {float(fx["dS_f64"].abs().max()):.3e}  ddustbin {float(fx["ddustbin_f64"]):.6f}  '
              f'f32-vs-f64 dS {float((fx["dS_f32"].double() - fx["dS_f64"]).abs().max()):.2e}')


if __name__ == '__main__':
    main()
