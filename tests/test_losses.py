"""Matching loss (SURVEY.md section 8, row f1 - the loss half): the oracle against outputs of the UNMODIFIED reference
``utils.losses.criterion`` and torch autograd through it (tests/golden/loss_*.pt, minted by oracle/gen_golden_loss.py), and the
CUDA path (og_criterion_fwd through openglue_b200.criterion / criterion_with_grad) against both.
Floating point: the loss is a mean of <= N + M log-scores; tolerance 2e-6 relative (fp32 summation order), gradients exact
to 1 ulp of their single division."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_oracle as L                                # noqa: E402  (checker only)
from oracle.gen_golden_loss import synthetic_labels                # noqa: E402  (label generator, no reference import at module level)

GOLDEN = ['loss_small', 'loss_medium', 'loss_empty']


def _fx(name):
    fx = torch.load(os.path.join(ROOT, 'tests', 'golden', name + '.pt'), weights_only=False)
    gt0, gt1, scores = synthetic_labels(*fx['case'])
    return fx, {'gt_matches0': gt0, 'gt_matches1': gt1}, scores


@pytest.mark.parametrize('name', GOLDEN)
def test_oracle_matches_reference(name):
    fx, y_true, scores = _fx(name)
    out = L.criterion(y_true, {'scores': scores.double()})
    assert abs(float(out['loss']) - float(fx['loss_f64'])) <= 1e-12
    assert float(out['metric_loss']) == 0.0 == float(fx['metric_loss_f64'])
    out32 = L.criterion(y_true, {'scores': scores})
    assert abs(float(out32['loss']) - float(fx['loss_f32'])) <= 2e-6 * abs(float(fx['loss_f32']))
    g = L.criterion_grad(y_true, tuple(scores.shape))
    assert (g - fx['dscores_f64'].to_dense()).abs().max() <= 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize('name', GOLDEN)
def test_cuda_criterion_matches_reference(name):
    from openglue_b200.losses import criterion, criterion_with_grad
    fx, y_true, scores = _fx(name)
    dev = 'cuda:0'
    yt = {k: v.to(dev) for k, v in y_true.items()}
    out = criterion(yt, {'scores': scores.to(dev)}, margin=None)
    assert abs(float(out['loss']) - float(fx['loss_f64'])) <= 2e-6 * abs(float(fx['loss_f64']))
    assert float(out['metric_loss']) == 0.0
    out2, ds = criterion_with_grad(yt, {'scores': scores.to(dev)}, grad_scale=0.75)
    assert float(out2['loss']) == float(out['loss'])                      # deterministic
    want = 0.75 * fx['dscores_f64'].to_dense()
    assert (ds.cpu().double() - want).abs().max() <= 2e-7 * float(want.abs().max())
    assert int((ds != 0).sum()) == fx['dscores_f64']._nnz()
    with pytest.raises(NotImplementedError):
        criterion(yt, {'scores': scores.to(dev)}, margin=0.5)


@pytest.mark.gpu
def test_cuda_criterion_on_the_path_outputs():
    """training_step order (matching_module.py:84-101): generate_gt_matches -> SuperGlue.forward -> criterion, all through the C ABI,
    against the oracles run on the same tensors."""
    from openglue_b200 import generate_gt_matches
    from openglue_b200.losses import criterion
    from openglue_b200.superglue import SuperGlue
    from openglue_b200.synthetic import default_config, synthetic_pairs, synthetic_state_dict
    from oracle import superglue_oracle as O
    from oracle import gt_matches_oracle as G
    dev = 'cuda:0'
    cfg = default_config(descriptor_dim=64, num_stages=2, num_iters=30)
    sd = synthetic_state_dict(cfg, seed=0)
    data = synthetic_pairs(3, 150, 170, 64, 1, family='planted', seed=5)
    H = torch.tensor([[0.9, 0.0, 20.0], [0.0, 0.9, 20.0], [0.0, 0.0, 1.0]]).repeat(3, 1, 1)     # the planted similarity of synthetic_pairs
    tf = {'type': ['perspective'] * 3, 'H': H}
    model = SuperGlue(dict(cfg)).eval()
    model.load_state_dict(sd)
    model = model.to(dev)
    f0 = {'keypoints': data['keypoints0'].to(dev), 'side_info': data['side_info0'].to(dev), 'local_descriptors': data['local_descriptors0'].to(dev)}
    f1 = {'keypoints': data['keypoints1'].to(dev), 'side_info': data['side_info1'].to(dev), 'local_descriptors': data['local_descriptors1'].to(dev)}
    batch = {'transformation': {'type': tf['type'], 'H': H.to(dev)}, 'image0_size': data['image0_size'], 'image1_size': data['image1_size']}
    d, y_true = generate_gt_matches(batch, f0, f1, 3.0, 5.0)
    y_pred = model(d)
    loss = criterion(y_true, y_pred)
    g0, g1, _ = G.gt_matches(data['keypoints0'], data['keypoints1'], tf)
    ref = O.run(sd, cfg, data, 0.2)
    want = L.criterion({'gt_matches0': g0, 'gt_matches1': g1}, {'scores': ref['scores'].double()})
    assert int((y_true['gt_matches0'].cpu() >= 0).sum()) > 200              # the planted correspondences are the labels
    assert abs(float(loss['loss']) - float(want['loss'])) <= 1e-4
