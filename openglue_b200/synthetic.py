"""Deterministic synthetic keypoint sets, descriptors and weights for the matching core.

Shapes and distributions follow SURVEY.md section 8(d): keypoints ~ U[0,W-1]x[0,H-1] pixels,
responses / side-info ~ U[0,1], descriptors ~ N(0,I) L2-normalised (SuperPoint / HardNet
descriptors are unit-norm).  Two input families:

* ``flat``    - two independent random sets (log-scores nearly uniform; worst case for ties),
* ``planted`` - a share of image-1 keypoints are a known permutation of image-0 keypoints
                (descriptor + small noise, keypoints mapped by a fixed similarity) and the
                descriptors are scaled so planted pairs clear ``match_threshold``.

Everything is generated on the CPU with an explicit ``torch.Generator`` so the same
tensors appear on every box; callers move them to the device (or pin them) themselves.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

IMAGE_WH = (960, 720)          # reference config/config_cached.yaml:12  target_size


def default_config(descriptor_dim: int = 256, num_stages: int = 9, num_heads: int = 4,
                   num_iters: int = 100, side_info_size: int = 1, reg: float = 1.0,
                   residual: bool = True, use_offset: bool = False,
                   hidden_layers_sizes=(32, 64, 128)) -> dict:
    """The nested dict the reference's SuperGlue(config) takes (superglue.py:12-27,
    back-filled the way matching_module.py:35-43 does)."""
    return {
        'descriptor_dim': descriptor_dim,
        'positional_encoding': {'output_size': descriptor_dim, 'side_info_size': side_info_size,
                                'hidden_layers_sizes': list(hidden_layers_sizes)},
        'attention_gnn': {'num_stages': num_stages, 'embed_dim': descriptor_dim,
                          'num_heads': num_heads, 'attention': 'softmax', 'use_offset': use_offset},
        'dustbin_score_init': 1.0,
        'otp': {'num_iters': num_iters, 'reg': reg},
        'residual': residual,
    }


# BASELINE.json configs (per-GPU batch, N, M, config kwargs)
BASELINE_CONFIGS = {
    'C1': dict(batch=1, n=512, m=512, cfg=dict(descriptor_dim=256, num_stages=9, num_iters=20)),
    'C2': dict(batch=32, n=1024, m=1024, cfg=dict(descriptor_dim=256, num_stages=9, num_iters=100)),
    'C3': dict(batch=16, n=2048, m=2048, cfg=dict(descriptor_dim=256, num_stages=9, num_iters=100)),
    'C4': dict(batch=32, n=2048, m=2048, cfg=dict(descriptor_dim=256, num_stages=9, num_iters=100)),
    'C5': dict(batch=1, n=4096, m=1024, cfg=dict(descriptor_dim=128, num_stages=18, num_iters=50,
                                                  side_info_size=6)),
}


def _conv_init(gen: torch.Generator, out_c: int, in_c: int):
    """PyTorch's default Conv1d(k=1) init: weight and bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    bound = 1.0 / math.sqrt(in_c)
    w = (torch.rand(out_c, in_c, 1, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_c, generator=gen) * 2 - 1) * bound
    return w, b


def _bn_init(gen: torch.Generator, c: int, randomize: bool):
    if not randomize:
        return {'weight': torch.ones(c), 'bias': torch.zeros(c), 'running_mean': torch.zeros(c),
                'running_var': torch.ones(c), 'num_batches_tracked': torch.tensor(0)}
    return {'weight': 0.5 + torch.rand(c, generator=gen),
            'bias': 0.2 * torch.randn(c, generator=gen),
            'running_mean': 0.1 * torch.randn(c, generator=gen),
            'running_var': 0.5 + torch.rand(c, generator=gen),
            'num_batches_tracked': torch.tensor(7)}


def synthetic_state_dict(config: dict, seed: int = 0, randomize_bn: bool = True,
                         randomize_scalars: bool = True) -> Dict[str, torch.Tensor]:
    """A state_dict with exactly the reference's keys and shapes,
    filled from a seeded generator.  ``randomize_bn`` gives BatchNorm non-trivial affine
    and running statistics so that BN folding is really exercised."""
    gen = torch.Generator().manual_seed(seed)
    d = config['descriptor_dim']
    sd: Dict[str, torch.Tensor] = {}
    if config.get('residual', False):
        sd['mix_coefs'] = (0.5 * torch.randn(d, 1, generator=gen)) if randomize_scalars else torch.zeros(d, 1)
    sd['dustbin_score'] = torch.tensor(float(config['dustbin_score_init']))
    pe = config['positional_encoding']
    sizes = [2 + pe['side_info_size'], *(pe.get('hidden_layers_sizes') or []), pe['output_size']]
    for i in range(1, len(sizes)):
        w, b = _conv_init(gen, sizes[i], sizes[i - 1])
        j = 3 * (i - 1)
        sd[f'positional_encoding.encoder.{j}.weight'] = w
        sd[f'positional_encoding.encoder.{j}.bias'] = b
        if i < len(sizes) - 1:
            for k, v in _bn_init(gen, sizes[i], randomize_bn).items():
                sd[f'positional_encoding.encoder.{j + 2}.{k}'] = v
    for layer in range(2 * config['attention_gnn']['num_stages']):
        p = f'attention_gnn.layers.{layer}.module.'
        for name in ('in_proj_q', 'in_proj_k', 'in_proj_v', 'out_proj'):
            w, b = _conv_init(gen, d, d)
            sd[p + f'mha.{name}.weight'], sd[p + f'mha.{name}.bias'] = w, b
        w, b = _conv_init(gen, 2 * d, 2 * d)
        sd[p + 'fc.0.weight'], sd[p + 'fc.0.bias'] = w, b
        for k, v in _bn_init(gen, 2 * d, randomize_bn).items():
            sd[p + f'fc.2.{k}'] = v
        w, b = _conv_init(gen, d, 2 * d)
        sd[p + 'fc.3.weight'], sd[p + 'fc.3.bias'] = w, b
    w, b = _conv_init(gen, d, d)
    sd['linear_proj.weight'], sd['linear_proj.bias'] = w, b
    return sd


def synthetic_pairs(batch: int, n: int, m: int, descriptor_dim: int, side_info_size: int = 1,
                    family: str = 'planted', seed: int = 1234, image_wh=IMAGE_WH,
                    shared: float = 0.7, noise: float = 0.03, desc_scale: Optional[float] = None
                    ) -> Dict[str, object]:
    """A batch of image pairs as the ``data`` dict SuperGlue.forward takes
    (superglue.py:30-38): keypoints{0,1} [B,n,2] pixels, side_info{0,1} [B,n,S],
    local_descriptors{0,1} [B,n,d] keypoint-major, image{0,1}_size = (W, H).
    For ``planted`` also returns ``planted_matches0`` [B,n] (index into image 1 or -1)."""
    gen = torch.Generator().manual_seed(seed)
    w, h = image_wh
    wh = torch.tensor([w - 1.0, h - 1.0])

    def unit(x):
        return x / x.norm(dim=-1, keepdim=True)

    k0 = torch.rand(batch, n, 2, generator=gen) * wh
    s0 = torch.rand(batch, n, side_info_size, generator=gen)
    d0 = unit(torch.randn(batch, n, descriptor_dim, generator=gen))
    k1 = torch.rand(batch, m, 2, generator=gen) * wh
    s1 = torch.rand(batch, m, side_info_size, generator=gen)
    d1 = unit(torch.randn(batch, m, descriptor_dim, generator=gen))
    planted = torch.full((batch, n), -1, dtype=torch.int64)
    if family == 'planted':
        if desc_scale is None:
            desc_scale = 32.0
        n_shared = int(shared * min(n, m))
        for b in range(batch):
            src = torch.randperm(n, generator=gen)[:n_shared]
            dst = torch.randperm(m, generator=gen)[:n_shared]
            d1[b, dst] = unit(d0[b, src] + noise * torch.randn(n_shared, descriptor_dim, generator=gen))
            k1[b, dst] = (0.9 * k0[b, src] + 20.0).clamp_(max=wh)
            s1[b, dst] = s0[b, src]
            planted[b, src] = dst
    elif family != 'flat':
        raise ValueError(f'unknown input family {family!r}')
    scale = 1.0 if desc_scale is None else float(desc_scale)
    return {
        'keypoints0': k0, 'keypoints1': k1, 'side_info0': s0, 'side_info1': s1,
        'local_descriptors0': d0 * scale, 'local_descriptors1': d1 * scale,
        'image0_size': (w, h), 'image1_size': (w, h), 'planted_matches0': planted,
    }


def synthetic_gt_scene(batch: int, n: int, m: int, kind: str = 'perspective', seed: int = 0, depth_image: bool = False,
                       width: int = 640, height: int = 480, planted: float = 0.6, noise: float = 0.4,
                       missing_depth: float = 0.1) -> dict:
    """Two keypoint sets related by a known transformation, in the layout ``generate_gt_matches`` consumes
    (reference models/gt_matches_generation.py:17-36: ``data['transformation']`` as collated by the datasets,
    ``features{0,1}['keypoints']`` [B, n, 2] in pixels).  A fraction ``planted`` of image-1 keypoints are noisy
    reprojections of image-0 keypoints; the rest are uniform.  kind: 'perspective' | '3d_reprojection'
    (per-keypoint depth [B, n], or depth images [B, height, width] when ``depth_image``)."""
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.rand(*s, generator=g)
    size = torch.tensor([width - 1.0, height - 1.0])
    k0 = rnd(batch, n, 2) * (size - 8.0) + 4.0
    n_pl = min(int(planted * m), n)
    tf = {'type': [kind] * batch}
    if kind == 'perspective':
        H = torch.eye(3).repeat(batch, 1, 1)
        H[:, :2, :2] += (rnd(batch, 2, 2) - 0.5) * 0.08
        H[:, :2, 2] += (rnd(batch, 2) - 0.5) * 30.0
        H[:, 2, :2] += (rnd(batch, 2) - 0.5) * 4e-5
        tf['H'] = H
        hom = torch.cat([k0, torch.ones(batch, n, 1)], 2) @ H.transpose(1, 2)
        proj = hom[..., :2] / hom[..., 2:3]
        z1_of_0 = None
    elif kind == '3d_reprojection':
        f0, f1 = 500.0 + 100.0 * rnd(batch), 520.0 + 100.0 * rnd(batch)
        K0, K1 = torch.zeros(batch, 3, 3), torch.zeros(batch, 3, 3)
        for K, f in ((K0, f0), (K1, f1)):
            K[:, 0, 0] = f; K[:, 1, 1] = f * 1.01; K[:, 0, 2] = width / 2; K[:, 1, 2] = height / 2; K[:, 2, 2] = 1.0
        ang = (rnd(batch, 3) - 0.5) * 0.12
        R = torch.linalg.matrix_exp(torch.stack([torch.stack([torch.zeros(batch), -ang[:, 2], ang[:, 1]], 1),
                                                 torch.stack([ang[:, 2], torch.zeros(batch), -ang[:, 0]], 1),
                                                 torch.stack([-ang[:, 1], ang[:, 0], torch.zeros(batch)], 1)], 1))
        T = (rnd(batch, 3) - 0.5) * torch.tensor([0.6, 0.4, 0.2]) + torch.tensor([0.0, 0.0, 0.15])
        if depth_image:
            yy, xx = torch.meshgrid(torch.arange(height).float(), torch.arange(width).float(), indexing='ij')
            base = 4.0 + 1.5 * torch.sin(xx / 90.0)[None] + 1.0 * torch.cos(yy / 70.0)[None] + rnd(batch, 1, 1)
            d0_img = base.clone()
            d0_img[rnd(batch, height, width) < missing_depth] = 0.0
            idx = k0.long()
            z0 = d0_img[torch.arange(batch)[:, None], idx[..., 1], idx[..., 0]]
        else:
            z0 = 2.0 + 6.0 * rnd(batch, n)
            z0[rnd(batch, n) < missing_depth] = 0.0
        rays = torch.cat([k0, torch.ones(batch, n, 1)], 2) @ torch.linalg.inv(K0).transpose(1, 2)
        X1 = (rays * torch.where(z0 > 0, z0, torch.ones_like(z0)).unsqueeze(-1)) @ R.transpose(1, 2) + T[:, None]
        p = X1 @ K1.transpose(1, 2)
        proj = p[..., :2] / p[..., 2:3]
        z1_of_0 = X1[..., 2]
        tf.update(K0=K0, K1=K1, R=R, T=T)
    else:
        raise ValueError(kind)
    k1 = rnd(batch, m, 2) * (size - 8.0) + 4.0
    perm = torch.stack([torch.randperm(n, generator=g)[:n_pl] for _ in range(batch)])
    slot = torch.stack([torch.randperm(m, generator=g)[:n_pl] for _ in range(batch)])
    bi = torch.arange(batch)[:, None]
    planted_xy = proj[bi, perm] + noise * torch.randn(batch, n_pl, 2, generator=g)
    inside = ((planted_xy > 2.0) & (planted_xy < size - 2.0)).all(-1)
    k1[bi, slot] = torch.where(inside.unsqueeze(-1), planted_xy, k1[bi, slot])
    if kind == '3d_reprojection':
        if depth_image:
            d1_img = 3.0 + 2.0 * rnd(batch, height, width)
            d1_img[rnd(batch, height, width) < missing_depth] = 0.0
            idx1 = k1.long()
            zz = torch.where(inside, z1_of_0[bi, perm], d1_img[bi, idx1[bi, slot][..., 1], idx1[bi, slot][..., 0]])
            d1_img[bi, idx1[bi, slot][..., 1], idx1[bi, slot][..., 0]] = zz     # consistent depth under the planted points
            tf['depth0'], tf['depth1'] = d0_img, d1_img
        else:
            z1 = 2.0 + 6.0 * rnd(batch, m)
            z1[bi, slot] = torch.where(inside, z1_of_0[bi, perm], z1[bi, slot])
            z1[rnd(batch, m) < missing_depth] = 0.0
            tf['depth0'], tf['depth1'] = z0, z1
    return {'keypoints0': k0.contiguous(), 'keypoints1': k1.contiguous(), 'transformation': tf}

