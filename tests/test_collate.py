"""Cached-feature loader + collation: the oracle against outputs of the UNMODIFIED reference
``MegaDepthPairsDataModuleFeatures.stack_keypoints_batch`` (tests/golden/collate_*.pt, oracle/gen_golden_collate.py), the on-disk
round trip of the feature store, and the CUDA collation (og_collate_fwd through openglue_b200.collate_features) against both.
Index / gather work: bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import collate_oracle as CO                             # noqa: E402  (checker only)
from oracle.gen_golden_collate import CASES, synthetic_items        # noqa: E402  (input generator; no reference import at module level)

GOLDEN = list(CASES)
KEYS = ('lafs0', 'lafs1', 'scores0', 'scores1', 'descriptors0', 'descriptors1')


def _fx(name):
    fx = torch.load(os.path.join(ROOT, 'tests', 'golden', name + '.pt'), weights_only=False)
    return fx, synthetic_items(fx['case'])


def _random_selection(fx, items, target):
    """the torch.randperm draws of the reference's loop (megadepth_datamodule.py:145-148), replayed from its seed"""
    torch.manual_seed(fx['rng_seed'])
    sel = torch.zeros(2 * len(items), target, dtype=torch.int32)
    for b, it in enumerate(items):
        for i in (0, 1):
            c = it[f'lafs{i}'].size(0)
            if c > target:
                sel[2 * b + i] = torch.randperm(c)[:target].to(torch.int32)
    return sel


def _same(got, want):
    for k in KEYS:
        assert torch.equal(got[k].cpu(), want[k]), k
    for k in ('depth0', 'depth1', 'K0', 'K1', 'R', 'T'):
        assert torch.equal(got['transformation'][k].cpu(), want['transformation'][k]), k
    assert got['transformation']['type'] == want['transformation']['type']
    assert tuple(got['image0_size']) == tuple(want['image0_size'])


@pytest.mark.parametrize('name', GOLDEN)
def test_oracle_matches_reference(name):
    fx, items = _fx(name)
    b, k, d, counts, rnd, seed = fx['case']
    sel = _random_selection(fx, items, k) if rnd else None
    _same(CO.stack_keypoints_batch(items, k, sel), fx['out'])


def test_feature_store_roundtrip(tmp_path):
    """extract_features.save_outputs' four arrays -> .npz -> FeatureStore: what the dataset's __getitem__ reads per image."""
    from openglue_b200.feature_cache import FeatureStore, save_features_npz
    g = torch.Generator().manual_seed(0)
    want = {}
    for n, c in (('img_a', 77), ('img_b', 1)):
        arrs = (torch.randn(c, 2, 3, generator=g).numpy(), torch.rand(c, generator=g).numpy(), torch.randn(c, 24, generator=g).numpy(), np.array([960, 720]))
        save_features_npz(str(tmp_path), n, *arrs)
        want[n] = arrs
    store = FeatureStore(str(tmp_path), pin=False)
    assert store.names() == ['img_a', 'img_b']
    for n, (lafs, sc, de, size) in want.items():
        it = store[n]
        assert np.array_equal(it['lafs'].numpy(), lafs) and np.array_equal(it['scores'].numpy(), sc) and np.array_equal(it['descriptors'].numpy(), de)
        assert it['size'] == (960, 720) and store[n] is it                       # cached: read once
    with pytest.raises(ImportError):                                            # the reference's .h5 form needs an HDF5 reader
        open(os.path.join(str(tmp_path), 'img_c_lafs.h5'), 'wb').close()
        FeatureStore(str(tmp_path), pin=False)['img_c']


@pytest.mark.gpu
@pytest.mark.parametrize('name', GOLDEN)
def test_cuda_collate_matches_reference(name):
    from openglue_b200.feature_cache import collate_features
    fx, items = _fx(name)
    b, k, d, counts, rnd, seed = fx['case']
    torch.manual_seed(fx['rng_seed'])                                           # random mode: the reference's generator state
    got = collate_features(items, k, random=rnd, device='cuda:0')
    torch.cuda.synchronize()
    _same(got, fx['out'])


@pytest.mark.gpu
def test_cuda_collate_feeds_the_path():
    """collate -> prepare_features_output's slicing -> generate_gt_matches -> SuperGlue, device-resident end to end."""
    from openglue_b200 import SuperGlue, generate_gt_matches
    from openglue_b200.feature_cache import collate_features
    from openglue_b200.synthetic import default_config, synthetic_state_dict
    fx, items = _fx('collate_topk')
    for it in items:                                                             # make it a geometric scene: identity pose, unit intrinsics
        it['transformation'].update(K0=torch.eye(3), K1=torch.eye(3), R=torch.eye(3), T=torch.zeros(3))
    out = collate_features(items, 64, device='cuda:0')
    cfg = default_config(descriptor_dim=32, num_stages=1, num_iters=5)
    model = SuperGlue(dict(cfg)).eval()
    model.load_state_dict(synthetic_state_dict(cfg, seed=0))
    model = model.to('cuda:0')
    feats = [{'keypoints': out[f'lafs{i}'][:, :, :, -1].contiguous(), 'side_info': out[f'scores{i}'].unsqueeze(-1),
              'local_descriptors': out[f'descriptors{i}']} for i in (0, 1)]            # models/features/utils.py:54-65
    data, y_true = generate_gt_matches({'transformation': out['transformation'], 'image0_size': out['image0_size'],
                                        'image1_size': out['image1_size']}, feats[0], feats[1], 3.0, 5.0)
    pred = model(data)
    assert pred['scores'].shape == (2, 65, 65) and torch.isfinite(pred['scores']).all()
    assert y_true['gt_matches0'].shape == (2, 64)
