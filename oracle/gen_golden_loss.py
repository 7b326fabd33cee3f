"""Mint golden vectors for the matching loss by running the UNMODIFIED reference ``utils.losses.criterion``
(ucuapps/OpenGlue @ /root/reference) and torch autograd through it.

TEST INFRASTRUCTURE.  Runs only in the build container; outputs are committed under tests/golden/loss_*.pt and pin
``oracle/loss_oracle.py`` (tests/test_losses.py).

    python oracle/gen_golden_loss.py
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
    # name: (batch, n, m, seed, matched share, ignore share, empty_sets)
    'loss_small':   (3, 37, 52, 1, 0.5, 0.1, False),
    'loss_medium':  (4, 300, 257, 2, 0.6, 0.05, False),
    'loss_empty':   (3, 20, 24, 3, 0.5, 0.1, True),        # pair 1 has no matched keypoint, pair 2 no unmatched one in image 1
}


def synthetic_labels(batch, n, m, seed, matched, ignore, empty_sets):
    """gt_matches0 / gt_matches1 with the reference's marks (models/gt_matches_generation.py:13-14: -1 unmatched, -2 ignore)
    and log-scores of a plausible magnitude; every matched pair is mutual, as generate_gt_matches produces them."""
    g = torch.Generator().manual_seed(seed)
    gt0 = torch.full((batch, n), -1, dtype=torch.int64)
    gt1 = torch.full((batch, m), -1, dtype=torch.int64)
    for b in range(batch):
        k = int(matched * min(n, m))
        if empty_sets and b == 1:
            k = 0
        src = torch.randperm(n, generator=g)[:k]
        dst = torch.randperm(m, generator=g)[:k]
        gt0[b, src] = dst
        gt1[b, dst] = src
        ig0 = torch.rand(n, generator=g) < ignore
        ig1 = torch.rand(m, generator=g) < ignore
        gt0[b, ig0 & (gt0[b] < 0)] = -2
        gt1[b, ig1 & (gt1[b] < 0)] = -2
        if empty_sets and b == 2:
            gt1[b, gt1[b] == -1] = -2
    scores = -8.0 * torch.rand(batch, n + 1, m + 1, generator=g) - 0.05
    return gt0, gt1, scores


def main():
    from utils.losses import criterion                                   # the reference, unmodified
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, case in CASES.items():
        gt0, gt1, scores = synthetic_labels(*case)
        b, n, m = case[:3]
        y_true = {'gt_matches0': gt0, 'gt_matches1': gt1}
        fx = {'case': case, 'reference': 'utils/losses.py:7-53 @ /root/reference, torch ' + torch.__version__}
        for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            s = scores.to(dtype).clone().requires_grad_(True)
            y_pred = {'context_descriptors0': torch.zeros(b, 4, n, dtype=dtype), 'context_descriptors1': torch.zeros(b, 4, m, dtype=dtype),
                      'scores': s}
            out = criterion(y_true, y_pred, margin=None)
            out['loss'].backward()
            fx[f'loss_{tag}'] = out['loss'].detach().clone()
            fx[f'metric_loss_{tag}'] = torch.as_tensor(out['metric_loss']).detach().clone()
            fx[f'dscores_{tag}'] = s.grad.detach().clone().to_sparse()
        torch.save(fx, os.path.join(out_dir, name + '.pt'))
        print(f'{name}: loss {float(fx["loss_f64"]):.6f}  metric_loss {float(fx["metric_loss_f64"])}  nnz(dscores) {fx["dscores_f64"]._nnz()}')


if __name__ == '__main__':
    main()
