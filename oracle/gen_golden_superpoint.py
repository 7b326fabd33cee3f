"""Mint golden vectors for the SuperPoint front-end by running the UNMODIFIED reference
``models.features.superpoint.model.SuperPointNet`` (ucuapps/OpenGlue @ /root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container; outputs are committed under tests/golden/superpoint_*.pt.

kornia is not installed here (and is stubbed for every other fixture script): the ONE kornia function on this path,
``kornia.geometry.subpix.nms2d`` (model.py:7,89), is restated below from kornia >= 0.6.1 (the reference's requirements.txt pin) -
``NonMaximaSuppression2d``: a one-hot "neighbour to channel" convolution of the replicate-padded map whose CENTRE channel is all
zero, a max over those channels, ``mask = x > max``, ``x * mask`` - and injected into the stub, so that everything else (conv stack,
softmax, pixel shuffle, threshold, nonzero, remove_borders, top_k_keypoints, grid_sample, min_stack) is the reference's own code.
That piece is therefore pinned to the published algorithm, not to an execution of kornia ("parity unpinned" for nms2d alone).

    python oracle/gen_golden_superpoint.py
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('OPENGLUE_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

from gen_golden import _stub_modules  # noqa: E402


def nms2d(x: torch.Tensor, kernel_size, mask_only: bool = False) -> torch.Tensor:
    """kornia.geometry.subpix.nms2d (kornia 0.6.x, kornia/geometry/subpix/nms.py: NonMaximaSuppression2d.forward)."""
    B, CH, H, W = x.size()
    ky, kx = kernel_size
    numel = ky * kx
    weight = torch.eye(numel, dtype=x.dtype)
    weight[numel // 2, numel // 2] = 0                                   # the centre's channel stays all zero
    kernel = weight.view(numel, 1, ky, kx)
    pad = [(kx - 1) // 2, (kx - 1) // 2, (ky - 1) // 2, (ky - 1) // 2]
    max_non_center = F.conv2d(F.pad(x, pad, mode='replicate'), kernel.repeat(CH, 1, 1, 1), stride=1, groups=CH
                              ).view(B, CH, -1, H, W).max(dim=2)[0]
    mask = x > max_non_center
    return mask if mask_only else x * mask.to(x.dtype)


def synthetic_superpoint_state_dict(seed: int, descriptor_dim: int = 256):
    """He-scaled random weights: activations stay O(1) through the stack, the detector logits spread enough for a textured heat map."""
    g = torch.Generator().manual_seed(seed)
    shapes = [('conv1a', 1, 64, 3), ('conv1b', 64, 64, 3), ('conv2a', 64, 64, 3), ('conv2b', 64, 64, 3), ('conv3a', 64, 128, 3),
              ('conv3b', 128, 128, 3), ('conv4a', 128, 128, 3), ('conv4b', 128, 128, 3), ('convPa', 128, 256, 3), ('convPb', 256, 65, 1),
              ('convDa', 128, 256, 3), ('convDb', 256, descriptor_dim, 1)]
    sd = {}
    for name, ci, co, k in shapes:
        gain = 4.0 if name == 'convPb' else 1.0
        sd[name + '.weight'] = torch.randn(co, ci, k, k, generator=g) * (gain * (2.0 / (ci * k * k)) ** 0.5)
        sd[name + '.bias'] = 0.05 * torch.randn(co, generator=g)
    return sd


def synthetic_superpoint_bn_state_dict(seed: int):
    """SuperPointNetBn (model.py:132-199): the convolutions of `synthetic_superpoint_state_dict` + BatchNorm2d parameters and running
    statistics that are far from the identity (scale 0.5 .. 1.5, shifts, variances 0.3 .. 2), so that a wrong fold cannot pass."""
    sd = synthetic_superpoint_state_dict(seed)
    g = torch.Generator().manual_seed(1000 + seed)
    for name, ch in [('bn1a', 64), ('bn1b', 64), ('bn2a', 64), ('bn2b', 64), ('bn3a', 128), ('bn3b', 128), ('bn4a', 128), ('bn4b', 128),
                     ('bnPa', 256), ('bnPb', 65), ('bnDa', 256), ('bnDb', 256)]:
        sd[name + '.weight'] = 0.5 + torch.rand(ch, generator=g)
        sd[name + '.bias'] = 0.2 * torch.randn(ch, generator=g)
        sd[name + '.running_mean'] = 0.3 * torch.randn(ch, generator=g)
        sd[name + '.running_var'] = 0.3 + 1.7 * torch.rand(ch, generator=g)
        sd[name + '.num_batches_tracked'] = torch.tensor(100)
    return sd


def synthetic_images(batch, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, 1, h, w, generator=g)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    for b in range(batch):                                               # a few smooth blobs on top of the noise
        for _ in range(6):
            cy, cx = float(torch.rand(1, generator=g)) * h, float(torch.rand(1, generator=g)) * w
            img[b, 0] += 0.8 * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 6.0 ** 2))
    return img.clamp(0, 2) / 2


CASES = {
    # name: (batch, H, W, max_keypoints, keypoint_threshold, seed)
    'superpoint_all':  (2, 96, 128, -1, 0.0, 1),          # every NMS survivor, row-major order; counts differ -> min_stack's top-k
    'superpoint_topk': (2, 120, 160, 150, 0.0, 2),        # top-k of both images
    'superpoint_thr':  (1, 64, 64, 400, 0.02, 3),         # threshold active, fewer survivors than max_keypoints: order kept
    'superpoint_bn':   (2, 96, 128, 200, 0.0, 4),         # SuperPointNetBn: BatchNorm2d after every convolution (eval mode)
}


def main():
    _stub_modules()
    import kornia.geometry.subpix as subpix                              # the stub module
    subpix.nms2d = nms2d
    from models.features.superpoint.model import SuperPointNet as RefSuperPoint, SuperPointNetBn as RefSuperPointBn   # the reference, unmodified
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    only = sys.argv[1:]
    for name, (batch, h, w, maxk, thr, seed) in CASES.items():
        if only and name not in only:
            continue
        if name.endswith('_bn'):
            model = RefSuperPointBn(max_keypoints=maxk, keypoint_threshold=thr)
            print(name, model.load_state_dict(synthetic_superpoint_bn_state_dict(seed), strict=True))
        else:
            model = RefSuperPoint(max_keypoints=maxk, keypoint_threshold=thr)
            print(name, model.load_state_dict(synthetic_superpoint_state_dict(seed), strict=True))
        model.eval()
        img = synthetic_images(batch, h, w, seed)
        fx = {'case': (batch, h, w, maxk, thr, seed), 'image': img,
              'reference': 'models/features/superpoint/model.py:61-129 @ /root/reference (kornia nms2d restated), torch ' + torch.__version__}
        for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            m = model.to(dtype)
            with torch.no_grad():
                desc_map, cell = m._forward_layers(img.to(dtype))
                b_, _, hc, wc = cell.shape
                heat = cell.permute(0, 2, 3, 1).reshape(b_, hc, wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(b_, hc * 8, wc * 8)
                fx[f'heat_{tag}'] = heat.float().clone()
                fx[f'desc_map_{tag}'] = desc_map.float().clone()
                if dtype == torch.float32:           # the reference's forward is fp32-only (keypoints are cast with .float(), model.py:108)
                    fx['lafs'], fx['scores'], fx['descriptors'] = [t.clone() for t in m(img)]
        torch.save(fx, os.path.join(out_dir, name + '.pt'))
        print(f'{name}: lafs {tuple(fx["lafs"].shape)}; max |heat32 - heat64| {float((fx["heat_f32"] - fx["heat_f64"]).abs().max()):.2e}; '
              f'max |desc_map32 - desc_map64| {float((fx["desc_map_f32"] - fx["desc_map_f64"]).abs().max()):.2e}')

if __name__ == '__main__':
    main()
