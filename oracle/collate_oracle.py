"""CPU oracle for the cached-feature collation  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates ``MegaDepthPairsDataModuleFeatures.stack_keypoints_batch`` (reference data/megadepth_datamodule.py:105-168) with the
selection passed in explicitly for the random mode.  Pinned by tests/golden/collate_*.pt, minted from the UNMODIFIED reference
function by oracle/gen_golden_collate.py (tests/test_collate.py)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch


def stack_keypoints_batch(batch: List[Dict[str, Any]], target: int, select: Optional[torch.Tensor] = None) -> Dict[str, Any]:
    """megadepth_datamodule.py:105-168.  ``select`` [2B, target]: the randperm selection of the random mode (row 2b + image);
    None = top confidence (torch.topk, :149-150)."""
    B, D = len(batch), batch[0]['descriptors0'].size(1)
    res = {f'{k}{i}': torch.zeros(B, target, *shape) for i in (0, 1) for k, shape in (('lafs', (2, 3)), ('scores', ()), ('descriptors', (D,)))}
    depth = {i: torch.zeros(B, target) for i in (0, 1)}
    for b, item in enumerate(batch):
        for i in (0, 1):
            lafs, sc, de = item[f'lafs{i}'], item[f'scores{i}'], item[f'descriptors{i}']
            n = lafs.size(0)
            if n > target:                                                           # :146-157
                idx = select[2 * b + i].long() if select is not None else torch.topk(sc, target, dim=0).indices
                lafs, sc, de, n = lafs[idx], sc[idx], de[idx], target
            res[f'lafs{i}'][b, :n], res[f'scores{i}'][b, :n], res[f'descriptors{i}'][b, :n] = lafs, sc, de
            dimg = item['transformation'][f'depth{i}']
            depth[i][b, :n] = dimg[lafs[:, 1, 2].long(), lafs[:, 0, 2].long()]       # :154-157, :164-167
    tf0 = batch[0]['transformation']
    res['transformation'] = {'type': ['3d_reprojection'], **{k: torch.stack([x['transformation'][k] for x in batch]) for k in ('K0', 'K1', 'R', 'T')},
                             'depth0': depth[0], 'depth1': depth[1]}
    res['image0_size'], res['image1_size'] = batch[0]['image0_size'], batch[0]['image1_size']
    return res
