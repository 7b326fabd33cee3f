"""Mint golden vectors for ground-truth match generation by running the UNMODIFIED reference function
``models.gt_matches_generation.generate_gt_matches`` (ucuapps/OpenGlue @ /root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container; outputs are committed under tests/golden/gt_*.pt and pin
``oracle/gt_matches_oracle.py`` (tests/test_gt_matches.py).

    python oracle/gen_golden_gt.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('OPENGLUE_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from openglue_b200.synthetic import synthetic_gt_scene  # noqa: E402

CASES = {
    # name: (batch, n, m, kind, depth_image, seed)
    'gt_perspective':   (2, 300, 280, 'perspective', False, 1),
    'gt_3d_keypoint':   (2, 257, 310, '3d_reprojection', False, 2),
    'gt_3d_depthimage': (1, 200, 190, '3d_reprojection', True, 3),
    'gt_tiny':          (3, 5, 3, 'perspective', False, 4),
}


def main():
    from models.gt_matches_generation import generate_gt_matches          # the reference, unmodified
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, (b, n, m, kind, dimg, seed) in CASES.items():
        sc = synthetic_gt_scene(b, n, m, kind, seed=seed, depth_image=dimg)
        feats0 = {'keypoints': sc['keypoints0'], 'local_descriptors': torch.zeros(b, n, 4), 'side_info': torch.zeros(b, n, 1)}
        feats1 = {'keypoints': sc['keypoints1'], 'local_descriptors': torch.zeros(b, m, 4), 'side_info': torch.zeros(b, m, 1)}
        data = {'transformation': sc['transformation']}
        with torch.no_grad():
            new_data, y_true = generate_gt_matches(data, feats0, feats1, positive_threshold=3.0, negative_threshold=5.0)
        assert new_data['keypoints0'] is sc['keypoints0']
        fx = {'case': (b, n, m, kind, dimg, seed), 'gt_matches0': y_true['gt_matches0'], 'gt_matches1': y_true['gt_matches1'],
              'reference': 'models/gt_matches_generation.py:17-93 @ /root/reference, torch ' + torch.__version__}
        torch.save(fx, os.path.join(out_dir, name + '.pt'))
        g0 = y_true['gt_matches0']
        print(f'{name}: matched {(g0 >= 0).sum().item()}, unmatched {(g0 == -1).sum().item()}, ignored {(g0 == -2).sum().item()}')


if __name__ == '__main__':
    main()
