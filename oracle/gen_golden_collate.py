"""Mint golden vectors for the cached-feature collation by running the UNMODIFIED reference
``MegaDepthPairsDataModuleFeatures.stack_keypoints_batch`` (ucuapps/OpenGlue @ /root/reference, data/megadepth_datamodule.py:105-168;
pytorch_lightning / deepdish / cv2 stubbed: not installed, not on the path of this static method).

TEST INFRASTRUCTURE.  Runs only in the build container; outputs: tests/golden/collate_*.pt (tests/test_collate.py).

    python oracle/gen_golden_collate.py
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
    # name: (batch, target keypoints, descriptor dim, raw keypoint counts per (pair, image), random mode, seed)
    'collate_topk':   (2, 64, 32, [(150, 90), (64, 300)], False, 1),
    'collate_pad':    (3, 128, 16, [(40, 128), (100, 7), (129, 1)], False, 2),
    'collate_random': (2, 50, 8, [(200, 30), (51, 500)], True, 3),
    'collate_large':  (1, 2048, 64, [(5000, 2049)], False, 4),
}
H, W = 120, 160


def synthetic_items(case):
    """Items as MegaDepthPairsDatasetFeatures.__getitem__ returns them (data/megadepth_dataset.py:262-282)."""
    b, k, d, counts, rnd, seed = case
    g = torch.Generator().manual_seed(seed)
    items = []
    for (c0, c1) in counts:
        it = {}
        for i, c in ((0, c0), (1, c1)):
            lafs = torch.randn(c, 2, 3, generator=g)
            lafs[:, 0, 2] = torch.rand(c, generator=g) * (W - 1)          # x
            lafs[:, 1, 2] = torch.rand(c, generator=g) * (H - 1)          # y
            it[f'lafs{i}'] = lafs
            # DISTINCT scores (torch.rand has 24-bit resolution: 5000 draws collide with probability ~0.5, and torch.topk does not
            # define the order of equal scores): a random permutation of c levels + a jitter smaller than the level spacing
            it[f'scores{i}'] = (torch.randperm(c, generator=g).float() + 0.5 * torch.rand(c, generator=g)) / c
            it[f'descriptors{i}'] = torch.randn(c, d, generator=g)
        it['transformation'] = {'type': '3d_reprojection', 'K0': torch.randn(3, 3, generator=g), 'K1': torch.randn(3, 3, generator=g),
                                'R': torch.randn(3, 3, generator=g), 'T': torch.randn(3, generator=g),
                                'depth0': torch.rand(H, W, generator=g) * 10, 'depth1': torch.rand(H, W, generator=g) * 10}
        it['image0_size'] = (W, H); it['image1_size'] = (W, H)
        items.append(it)
    return items


def main():
    from oracle.gen_golden import _stub_modules
    _stub_modules()
    from data.megadepth_datamodule import MegaDepthPairsDataModuleFeatures          # the reference, unmodified
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, case in CASES.items():
        items = synthetic_items(case)
        b, k, d, counts, rnd, seed = case
        torch.manual_seed(1000 + seed)                                               # the reference draws from the global generator
        ref = MegaDepthPairsDataModuleFeatures.stack_keypoints_batch(items, k, random=rnd)
        fx = {'case': case, 'out': ref, 'rng_seed': 1000 + seed,
              'reference': 'data/megadepth_datamodule.py:105-168 @ /root/reference, torch ' + torch.__version__}
        torch.save(fx, os.path.join(out_dir, name + '.pt'))
        print(name, {kk: tuple(v.shape) for kk, v in ref.items() if torch.is_tensor(v)})


if __name__ == '__main__':
    main()
