"""Mint golden vectors for the TRAINING step by running the UNMODIFIED reference in ``train()``
mode: ``models.superglue.superglue.SuperGlue`` + ``utils.losses.criterion`` + torch autograd, i.e. what
``MatchingTrainingModule.training_step`` (models/matching_module.py:99-105) differentiates.

TEST INFRASTRUCTURE.  Runs only in the build container; outputs are committed under tests/golden/train_*.pt and are what
tests/test_training.py compares the CUDA training path with.

    python oracle/gen_golden_train.py

For every case: the reference module is built, ``load_state_dict(strict=True)`` of the synthetic weights (BatchNorm affine and
running buffers perturbed so that nothing is at its trivial initial value), one forward + backward in fp32 and in fp64, and the
fixture keeps: inputs, labels, the weights' seed + the perturbed running buffers, scores, context descriptors, loss, the gradient of
EVERY parameter and of the local descriptors (fp64 run, stored as fp32), and the BatchNorm running buffers after the step.
"""
from __future__ import annotations

import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('OPENGLUE_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

from openglue_b200.synthetic import default_config, synthetic_pairs, synthetic_state_dict  # noqa: E402
from gen_golden import _stub_modules  # noqa: E402
from gen_golden_loss import synthetic_labels  # noqa: E402

CASES = {
    # name: (batch, n, m, config overrides, data seed)
    'train_small':  (2, 50, 70, dict(descriptor_dim=64, num_stages=2, num_iters=10), 11),                       # head_dim 16
    'train_offset': (2, 64, 48, dict(descriptor_dim=128, num_heads=2, num_stages=1, num_iters=5, use_offset=True,
                                     residual=False, side_info_size=6), 12),                                    # head_dim 64
    'train_ragged': (3, 96, 131, dict(descriptor_dim=64, num_heads=2, num_stages=2, num_iters=20, reg=0.5), 13),  # head_dim 32, m % 4 != 0
}


def perturb_bn(sd, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        v = v.clone()
        if k.endswith('running_mean'):
            v = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('running_var'):
            v = 0.5 + torch.rand(v.shape, generator=g)
        out[k] = v
    return out


def main():
    _stub_modules()
    from models.superglue.superglue import SuperGlue as RefSuperGlue          # the reference, unmodified
    from utils.losses import criterion
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, (batch, n, m, kw, seed) in CASES.items():
        cfg = default_config(**kw)
        sd = perturb_bn(synthetic_state_dict(cfg, seed=seed), seed)
        data = synthetic_pairs(batch, n, m, cfg['descriptor_dim'], cfg['positional_encoding']['side_info_size'], family='planted', seed=seed)
        gt0, gt1, _ = synthetic_labels(batch, n, m, seed, 0.5, 0.1, False)
        fx = {'config': cfg, 'weights_seed': seed, 'bn_buffers': {k: v for k, v in sd.items() if 'running_' in k},
              'data': data, 'gt_matches0': gt0, 'gt_matches1': gt1,
              'reference': 'models/superglue/superglue.py + utils/losses.py @ /root/reference, train() mode, torch ' + torch.__version__}
        for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            model = RefSuperGlue(copy.deepcopy(cfg))
            print(name, tag, model.load_state_dict(sd, strict=True))
            model = model.to(dtype).train()
            d = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()}
            d['local_descriptors0'] = d['local_descriptors0'].clone().requires_grad_(True)
            d['local_descriptors1'] = d['local_descriptors1'].clone().requires_grad_(True)
            y_pred = model(d)
            loss = criterion({'gt_matches0': gt0, 'gt_matches1': gt1}, y_pred, margin=None)['loss']
            loss.backward()
            fx[f'loss_{tag}'] = loss.detach().clone()
            fx[f'scores_{tag}'] = y_pred['scores'].detach().clone()
            fx[f'context_descriptors0_{tag}'] = y_pred['context_descriptors0'].detach().clone()
            fx[f'grads_{tag}'] = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            fx[f'dlocal_descriptors0_{tag}'] = d['local_descriptors0'].grad.detach().clone()
            fx[f'dlocal_descriptors1_{tag}'] = d['local_descriptors1'].grad.detach().clone()
            fx[f'buffers_after_{tag}'] = {k: v.detach().clone() for k, v in model.named_buffers()}
        gn = sum(float(v.double().pow(2).sum()) for v in fx['grads_f64'].values()) ** 0.5
        err = max(float((fx['grads_f32'][k].double() - fx['grads_f64'][k].double()).abs().max()) for k in fx['grads_f32'])
        # keep the committed file small: the fp64 gradients rounded to fp32 + how far the reference's own fp32 run is from them
        fx['grads_f32_vs_f64_max_abs'] = err
        fx['grads'] = {k: v.float() for k, v in fx.pop('grads_f64').items()}
        fx.pop('grads_f32')
        fx['buffers_after'] = {k: v.float() for k, v in fx.pop('buffers_after_f64').items()}
        fx.pop('buffers_after_f32')
        fx.pop('context_descriptors0_f32')
        torch.save(fx, os.path.join(out_dir, name + '.pt'))
        print(f'{name}: loss {float(fx["loss_f64"]):.6f}  |grad| {gn:.4e}  max |g32 - g64| {err:.3e}  params {len(fx["grads"])}')


if __name__ == '__main__':
    main()
