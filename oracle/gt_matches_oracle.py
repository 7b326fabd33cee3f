"""CPU oracle for ground-truth match generation  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates ``generate_gt_matches`` (reference models/gt_matches_generation.py:17-93) and the reprojection helpers
it calls (utils/misc.py:21-103) with the same ATen ops (matmul, torch.linalg.inv, torch.cdist, min, gather), i.e.
the step that runs immediately before the matching core in the reference's training / validation step
(models/matching_module.py:84-93).  Only ``tests/`` import it.

EFFECTIVE semantics.  The reference's threshold refinements (gt_matches_generation.py:56-67) and its last two
lines (:76-78) assign through ``tensor[mask][cond] = v`` - boolean-mask indexing returns a copy, so these
statements change nothing.  What the function really returns is
    gt_matches0[i] = j   if j = argmin_j' |T(k0_i) - k1_j'| and i = argmin_i' |T^-1(k1_j) - k0_i'|   (mutual NN)
                   = -1  otherwise                                                        (UNMATCHED_INDEX)
                   = -2  where the reprojection mask of k0_i is False (unknown depth)      (IGNORE_INDEX, :72-73)
and symmetrically gt_matches1; ``positive_threshold`` / ``negative_threshold`` have no effect.  The oracle
restates exactly that and ``tests/test_gt_matches.py`` pins it to outputs of the unmodified reference function
(``oracle/gen_golden_gt.py`` -> ``tests/golden/gt_*.pt``).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

UNMATCHED_INDEX = -1      # gt_matches_generation.py:13
IGNORE_INDEX = -2         # gt_matches_generation.py:14


def perspective_transform(kpts: torch.Tensor, H: torch.Tensor, eps: float = 1e-8):
    """utils/misc.py:62-71"""
    b, n, _ = kpts.shape
    hom = torch.cat([kpts, torch.ones(b, n, 1, dtype=kpts.dtype)], dim=2)
    out = torch.matmul(hom, H.transpose(1, 2).contiguous())
    out = out[..., :2] / (out[..., 2].unsqueeze(-1) + eps)
    return out, torch.ones(b, n, dtype=torch.bool)


def reproject_3d(kpts: torch.Tensor, K0, K1, T, R, depth0, eps: float = 1e-8):
    """utils/misc.py:74-103"""
    b, n, _ = kpts.shape
    hom = torch.cat([kpts, torch.ones(b, n, 1, dtype=kpts.dtype)], dim=2)
    rays = torch.matmul(hom, torch.linalg.inv(K0).transpose(1, 2).contiguous())
    if depth0.dim() == 2:
        depth = depth0
    else:                                                    # depth image: nearest-lower pixel of the keypoint (:90-97)
        idx = kpts.type(torch.int64)
        depth = depth0[torch.arange(b).unsqueeze(-1), idx[..., 1], idx[..., 0]]
    mask = ~torch.isclose(depth, depth.new_tensor(0.0))
    x = rays * depth.unsqueeze(-1)
    x = torch.matmul(x, R.transpose(1, 2).contiguous()) + T.unsqueeze(1)
    x = torch.matmul(x, K1.transpose(1, 2).contiguous())
    return x[..., :2] / (x[..., 2].unsqueeze(-1) + eps), mask


def reproject_keypoints(kpts, tf):
    """utils/misc.py:21-34"""
    if tf['type'][0] == 'perspective':
        return perspective_transform(kpts, tf['H'])
    if tf['type'][0] == '3d_reprojection':
        return reproject_3d(kpts, tf['K0'], tf['K1'], tf['T'], tf['R'], tf['depth0'])
    raise ValueError(f"Unknown transformation type {tf['type'][0]}.")


def inverse_transformation(tf):
    """utils/misc.py:37-59"""
    if tf['type'][0] == 'perspective':
        return {'type': tf['type'], 'H': torch.linalg.inv(tf['H'])}
    if tf['type'][0] == '3d_reprojection':
        r_t = tf['R'].transpose(1, 2).contiguous()
        return {'type': tf['type'], 'K0': tf['K1'], 'K1': tf['K0'], 'R': r_t,
                'T': -torch.matmul(r_t, tf['T'].unsqueeze(-1)).squeeze(-1),
                'depth0': tf['depth1'], 'depth1': tf['depth0']}
    raise ValueError(f"Unknown transformation type {tf['type'][0]}.")


def gt_matches(kpts0: torch.Tensor, kpts1: torch.Tensor, transformation: dict) -> Tuple[torch.Tensor, torch.Tensor, Dict]:
    """gt_matches_generation.py:29-78 (effective semantics, see the module docstring).
    Returns (gt_matches0 [B,N] int64, gt_matches1 [B,M] int64, extras for the tests)."""
    inv = inverse_transformation(transformation)
    n, m = kpts0.shape[1], kpts1.shape[1]
    k0t, mask0 = reproject_keypoints(kpts0, transformation)                     # :38
    k1t, mask1 = reproject_keypoints(kpts1, inv)                                # :39
    d01 = torch.cdist(k0t, kpts1, p=2)                                          # :40
    d10 = torch.cdist(k1t, kpts0, p=2)                                          # :41
    _, nn0 = d01.min(2)                                                         # :43
    _, nn1 = d10.min(2)                                                         # :44
    gt0, gt1 = nn0.clone(), nn1.clone()
    ok0 = torch.arange(n).unsqueeze(0) == gt1.gather(1, gt0)                    # :47
    gt0[~ok0] = UNMATCHED_INDEX                                                 # :48
    ok1 = torch.arange(m).unsqueeze(0) == gt0.gather(1, gt1)                    # :50 (gt0 already carries the -1s)
    gt1[~ok1] = UNMATCHED_INDEX                                                 # :51
    gt0[~mask0] = IGNORE_INDEX                                                  # :72
    gt1[~mask1] = IGNORE_INDEX                                                  # :73
    return gt0, gt1, {'kpts0_transformed': k0t, 'kpts1_transformed': k1t, 'mask0': mask0, 'mask1': mask1,
                      'nn0': nn0, 'nn1': nn1}
