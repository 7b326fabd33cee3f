"""Mint golden vectors by running the UNMODIFIED reference (ucuapps/OpenGlue @ /root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference does not travel to
the GPU box); its outputs are committed under tests/golden/ and are what pins the oracle.

    python oracle/gen_golden.py            # rewrites tests/golden/*.pt

For every case it
  1. builds the reference ``models.superglue.superglue.SuperGlue(config).eval()``,
  2. ``load_state_dict(strict=True)`` of ``openglue_b200.synthetic.synthetic_state_dict`` (this
     also proves the key/shape contract of SURVEY.md section 8b),
  3. runs ``forward`` in fp32 and fp64 under no_grad,
  4. runs the reference's own ``MatchingTrainingModule.forward`` (models/matching_module.py:149-187,
     lightning/kornia/torchmetrics stubbed - they are not installed and are not on the path)
     on the same pair as a cached-features batch to get matches0 / matching_scores0,
and stores inputs (seed-reproducible, but stored for the small cases), weights seed and outputs.
"""
from __future__ import annotations

import copy
import os
import sys
import types
from unittest import mock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('OPENGLUE_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from openglue_b200.synthetic import default_config, synthetic_pairs, synthetic_state_dict  # noqa: E402


class _StubFinder:
    """Import hook: any module under these absent third-party roots resolves to a MagicMock
    (pytorch_lightning / kornia / torchmetrics ... are not installed; none is on the hot path)."""
    ROOTS = ('pytorch_lightning', 'kornia', 'kornia_moons', 'torchmetrics', 'cv2', 'deepdish',
             'albumentations', 'shutup', 'omegaconf', 'matplotlib', 'wandb')

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        if fullname.split('.')[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        m.__loader__ = self
        if spec.name == 'pytorch_lightning':
            m.LightningModule = torch.nn.Module
            m.LightningDataModule = type('LightningDataModule', (), {})
        if spec.name == 'torchmetrics':
            m.Metric = type('Metric', (torch.nn.Module,), {})
        if spec.name == 'kornia.feature.laf':
            m.LAFOrienter = type('LAFOrienter', (torch.nn.Module,), {})
        return m

    def exec_module(self, module):
        pass


def _stub_modules():
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())


CASES = {
    # name: (batch, n, m, config kwargs, family, store_full)
    'tiny_flat':      (2, 48, 40, dict(descriptor_dim=32, num_stages=2, num_iters=10), 'flat', True),
    'tiny_planted':   (2, 64, 64, dict(descriptor_dim=64, num_stages=2, num_iters=25), 'planted', True),
    'tiny_offset_s6': (1, 40, 56, dict(descriptor_dim=32, num_stages=1, num_iters=7, side_info_size=6,
                                       use_offset=True, residual=False, reg=0.7), 'flat', True),
    'small_planted':  (1, 256, 200, dict(descriptor_dim=128, num_stages=3, num_iters=50), 'planted', True),
    # BASELINE.json configs[0] (the reference's CPU-runnable case): outputs stored sub-sampled
    'C1_planted':     (1, 512, 512, dict(descriptor_dim=256, num_stages=9, num_iters=20), 'planted', False),
    'C1_flat':        (1, 512, 512, dict(descriptor_dim=256, num_stages=9, num_iters=20), 'flat', False),
    # BASELINE.json configs[1], [2] (headline), [4] at full depth and full batch: exactly the tensors bench.py times on rank 0
    # (synthetic_pairs(batch, ..., seed=1234)).  matches0 / matching_scores0 for EVERY pair of the batch (the reference's
    # MatchingTrainingModule.forward in fp32); log-scores (fp32 and fp64 reference runs) for the first SCORED_PAIRS pairs.
    'C2_planted':     (32, 1024, 1024, dict(descriptor_dim=256, num_stages=9, num_iters=100), 'planted', False),
    'C3_planted':     (16, 2048, 2048, dict(descriptor_dim=256, num_stages=9, num_iters=100), 'planted', False),
    'C4_planted':     (32, 2048, 2048, dict(descriptor_dim=256, num_stages=9, num_iters=100), 'planted', False),   # configs[3]: 32 pairs / GPU
    'C3_flat':        (1, 2048, 2048, dict(descriptor_dim=256, num_stages=9, num_iters=100), 'flat', False),
    'C5_planted':     (1, 4096, 1024, dict(descriptor_dim=128, num_stages=18, num_iters=50, side_info_size=6), 'planted', False),
}
SCORED_PAIRS = 2               # pairs of a big batch whose log-scores are stored (sub-sampled)
MATCH_THRESHOLD = 0.2          # reference config/config.yaml:40


def run_reference(name):
    from models.superglue.superglue import SuperGlue  # the reference, unmodified
    batch, n, m, kw, family, full = CASES[name]
    cfg = default_config(**kw)
    sd = synthetic_state_dict(cfg, seed=0)
    data = synthetic_pairs(batch, n, m, cfg['descriptor_dim'], cfg['positional_encoding']['side_info_size'],
                           family=family, seed=1234)
    out = {}
    scored = min(batch, SCORED_PAIRS)
    data_scored = {k: (v[:scored] if torch.is_tensor(v) else v) for k, v in data.items()}
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        model = SuperGlue(copy.deepcopy(cfg)).eval()
        missing = model.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        model = model.to(dtype)
        d = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data_scored.items()}
        with torch.no_grad():
            res = model(d)
        out[tag] = {k: v.clone() for k, v in res.items()}

    # the reference's own post-processing (matching_module.py:149-187) on a cached-features batch
    _stub_modules()
    import models.matching_module as mmod
    laf_conv = types.SimpleNamespace(side_info_dim=cfg['positional_encoding']['side_info_size'] - 1)

    def fake_prepare(lafs, responses, desc, laf_converter, permute_desc=False, log_response=False):
        # cached-features path: hand the synthetic keypoints/side-info/descriptors through untouched
        return {'keypoints': lafs, 'side_info': responses, 'local_descriptors': desc}

    with mock.patch.object(mmod, 'get_laf_to_sideinfo_converter', lambda name: laf_conv), \
            mock.patch.object(mmod, 'get_augmentation_transform', lambda c: None), \
            mock.patch.object(mmod, 'prepare_features_output', fake_prepare):
        train_cfg = {'use_cached_features': True, 'augmentations': {'name': 'none'}, 'evaluation': False,
                     'match_threshold': MATCH_THRESHOLD}
        sg_cfg = copy.deepcopy(cfg)
        sg_cfg['laf_to_sideinfo_method'] = 'none'
        module = mmod.MatchingTrainingModule(train_cfg, {'descriptor_dim': cfg['descriptor_dim']}, sg_cfg)
        module.superglue.load_state_dict(sd, strict=True)
        module.eval()
        batch_in = {'lafs0': data['keypoints0'], 'scores0': data['side_info0'], 'descriptors0': data['local_descriptors0'],
                    'lafs1': data['keypoints1'], 'scores1': data['side_info1'], 'descriptors1': data['local_descriptors1'],
                    'image0_size': data['image0_size'], 'image1_size': data['image1_size']}
        with torch.no_grad():
            pred = mmod.MatchingTrainingModule.forward(module, batch_in)
    matches0, mscores0 = pred['matches0'], pred['matching_scores0']

    fx = {'name': name, 'config': cfg, 'weights_seed': 0, 'inputs_seed': 1234, 'family': family,
          'batch': batch, 'n': n, 'm': m, 'match_threshold': MATCH_THRESHOLD, 'scored_pairs': scored,
          'matches0': matches0, 'matching_scores0': mscores0,
          'ref32_vs_ref64_max_abs': float((out['f32']['scores'].double() - out['f64']['scores']).abs().max())}
    if full:
        fx['data'] = {k: v for k, v in data.items()}
        if sum(v.numel() for v in sd.values()) < 200_000:      # else regenerate from weights_seed
            fx['state_dict'] = sd
        fx['scores_f32'] = out['f32']['scores']
        fx['scores_f64'] = out['f64']['scores']
        fx['context_descriptors0_f32'] = out['f32']['context_descriptors0']
        fx['context_descriptors1_f32'] = out['f32']['context_descriptors1']
    else:
        # big case: inputs/weights are regenerated from the seeds; keep a strided sample + checksums
        s32, s64 = out['f32']['scores'], out['f64']['scores']
        sr, sc = (7, 5) if n <= 512 else (11, 13)
        fx['sample_stride'] = (sr, sc)
        fx['scores_f32_sample'] = s32[:, ::sr, ::sc].clone()
        fx['scores_f64_sample'] = s64[:, ::sr, ::sc].clone()
        fx['scores_f32_lastrow'] = s32[:, -1, :].clone()
        fx['scores_f32_lastcol'] = s32[:, :, -1].clone()
        fx['scores_f64_rowsum'] = s64.sum(2)
        fx['scores_f64_colsum'] = s64.sum(1)
        fx['ctx0_f32_sample'] = out['f32']['context_descriptors0'][:, ::4, ::8].clone()
        fx['ctx1_f32_sample'] = out['f32']['context_descriptors1'][:, ::4, ::8].clone()
        fx['row_argmax_f64'] = s64[:, :-1, :-1].argmax(2)
        top2 = s64[:, :-1, :-1].topk(2, dim=2).values
        fx['row_top2_gap_f64'] = (top2[..., 0] - top2[..., 1]).float()
        fx['col_argmax_f64'] = s64[:, :-1, :-1].argmax(1)
        top2c = s64[:, :-1, :-1].topk(2, dim=1).values
        fx['col_top2_gap_f64'] = (top2c[:, 0] - top2c[:, 1]).float()
        fx['matching_scores0_f64'] = s64[:, :-1, :-1].max(2).values.exp().float()      # before the mutual mask
    return fx


def main():
    outdir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        fx = run_reference(name)
        path = os.path.join(outdir, f'{name}.pt')
        torch.save(fx, path)
        nm = int((fx['matches0'] >= 0).sum())
        print(f'{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)  matches={nm}  '
              f'ref32-vs-ref64={fx["ref32_vs_ref64_max_abs"]:.2e}')


if __name__ == '__main__':
    main()
