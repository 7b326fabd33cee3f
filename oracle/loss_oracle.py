"""CPU oracle for the matching loss  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates ``criterion`` (reference utils/losses.py:7-53) for ``margin=None`` as plain torch, plus the analytic gradient of
'loss' with respect to the log-scores.  Pinned by tests/golden/loss_*.pt, minted by oracle/gen_golden_loss.py from the
UNMODIFIED reference function and torch autograd (tests/test_losses.py).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

Tensor = torch.Tensor


def _mean_weights(batch_idx: Tensor) -> Tensor:
    """utils/losses.py:18-19: 1 / (#entries of the same pair), via unique_consecutive on the (sorted) batch index."""
    _, inv, counts = torch.unique_consecutive(batch_idx, return_inverse=True, return_counts=True)
    return (1 / counts)[inv]


def criterion(y_true: Dict[str, Tensor], y_pred: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """reference utils/losses.py:7-53 with margin = None (metric terms are tensor(0): :56-58, :83-85)."""
    gt0, gt1, scores = y_true['gt_matches0'], y_true['gt_matches1'], y_pred['scores']
    b, i0 = torch.where(gt0 >= 0)                                        # :16-21 matched keypoints
    matched = (-scores[b, i0, gt0[b, i0]] * _mean_weights(b)).sum()
    b, i0 = torch.where(gt0 == -1)                                       # :29-33 unmatched in image 0 -> dustbin column
    un0 = (-scores[b, i0, -1] * _mean_weights(b)).sum()
    b, i1 = torch.where(gt1 == -1)                                       # :40-44 unmatched in image 1 -> dustbin row
    un1 = (-scores[b, -1, i1] * _mean_weights(b)).sum()
    zero = torch.zeros((), dtype=scores.dtype)
    return {'loss': (matched + 0.5 * (un0 + un1)) / scores.size(0), 'metric_loss': zero}     # :50-53


def criterion_grad(y_true: Dict[str, Tensor], scores_shape: Tuple[int, int, int], dtype=torch.float64) -> Tensor:
    """d loss / d scores: the scatter of the gather's weights."""
    gt0, gt1 = y_true['gt_matches0'], y_true['gt_matches1']
    B = scores_shape[0]
    g = torch.zeros(scores_shape, dtype=dtype)
    b, i0 = torch.where(gt0 >= 0)
    g[b, i0, gt0[b, i0]] = -_mean_weights(b).to(dtype) / B
    b, i0 = torch.where(gt0 == -1)
    g[b, i0, -1] = -0.5 * _mean_weights(b).to(dtype) / B
    b, i1 = torch.where(gt1 == -1)
    g[b, -1, i1] = -0.5 * _mean_weights(b).to(dtype) / B
    return g
