"""TEST INFRASTRUCTURE: a torch emulation of the schedule og_superglue_forward runs on the
packed (folded) weights, keypoint-major activations, fp64-capable.  It lets the CPU suite
check the host-side folding/packing logic (openglue_b200/packing.py) against the oracle
without a GPU.  Not used by the product."""
import math

import torch

from openglue_b200 import _cabi


def _get(packed, lib, cfg, tid, idx, shape):
    off = lib.og_packed_offset(cfg, tid, idx)
    n = 1
    for s in shape:
        n *= s
    return packed[off:off + n].reshape(shape)


def forward_packed(packed, config, cfg, data, dtype=torch.float64):
    lib = _cabi.lib()
    packed = packed.to(dtype)
    d, H = cfg.descriptor_dim, cfg.num_heads
    dh = d // H
    sizes = [2 + cfg.side_info_size] + [cfg.hidden[i] for i in range(cfg.num_hidden)] + [d]

    def kenc(kpts, side, wh):
        w, h = wh
        x = torch.cat([2 * kpts / torch.tensor([w - 1, h - 1], dtype=dtype) - 1, side], -1)
        for i in range(len(sizes) - 1):
            W = _get(packed, lib, cfg, _cabi.OG_T_KENC_W, i, (sizes[i + 1], sizes[i]))
            b = _get(packed, lib, cfg, _cabi.OG_T_KENC_B, i, (sizes[i + 1],))
            x = x @ W.T + b
            if i < len(sizes) - 2:
                x = torch.relu(x)
        return x

    def attn(q, k, v):
        B, nq, _ = q.shape
        q = q.view(B, nq, H, dh).transpose(1, 2)
        k = k.view(B, -1, H, dh).transpose(1, 2)
        v = v.view(B, -1, H, dh).transpose(1, 2)
        p = ((q @ k.transpose(-1, -2)) * dh ** -0.5).softmax(-1)
        return (p @ v).transpose(1, 2).reshape(B, nq, d)

    def prop(l, xq, xs):
        Wqkv = _get(packed, lib, cfg, _cabi.OG_T_QKV_W, l, (3 * d, d))
        bqkv = _get(packed, lib, cfg, _cabi.OG_T_QKV_B, l, (3 * d,))
        q = xq @ Wqkv[:d].T + bqkv[:d]
        kv = xs @ Wqkv[d:].T + bqkv[d:]
        o = attn(q, kv[..., :d], kv[..., d:])
        W1 = _get(packed, lib, cfg, _cabi.OG_T_FC1_W, l, (2 * d, 2 * d))
        b1 = _get(packed, lib, cfg, _cabi.OG_T_FC1_B, l, (2 * d,))
        W2 = _get(packed, lib, cfg, _cabi.OG_T_FC2_W, l, (d, 2 * d))
        b2 = _get(packed, lib, cfg, _cabi.OG_T_FC2_B, l, (d,))
        hid = torch.relu(torch.cat([xq, o], -1) @ W1.T + b1)
        return xq + hid @ W2.T + b2

    dt = lambda t: t.to(dtype)
    x0 = dt(data['local_descriptors0']) + kenc(dt(data['keypoints0']), dt(data['side_info0']), data['image0_size'])
    x1 = dt(data['local_descriptors1']) + kenc(dt(data['keypoints1']), dt(data['side_info1']), data['image1_size'])
    for l in range(cfg.num_layers):
        if l % 2 == 0:
            x0, x1 = prop(l, x0, x0), prop(l, x1, x1)
        else:
            x0 = prop(l, x0, x1)
            x1 = prop(l, x1, x0)
    Wp = _get(packed, lib, cfg, _cabi.OG_T_PROJ_W, 0, (d, d))
    bp = _get(packed, lib, cfg, _cabi.OG_T_PROJ_B, 0, (d,))
    rm = _get(packed, lib, cfg, _cabi.OG_T_PROJ_RMIX, 0, (d,))
    g0 = x0 @ Wp.T + bp + rm * dt(data['local_descriptors0'])
    g1 = x1 @ Wp.T + bp + rm * dt(data['local_descriptors1'])
    s = (g0 @ g1.transpose(1, 2)) * d ** -0.5
    dust = _get(packed, lib, cfg, _cabi.OG_T_DUSTBIN, 0, (1,))[0]
    # sinkhorn in the kernel's one-exp-per-element formulation (csrc/sinkhorn.cuh)
    B, n, m = s.shape
    z = torch.full((B, n + 1, m + 1), float(dust), dtype=dtype)
    z[:, :n, :m] = s
    z = z / cfg.sinkhorn_reg
    norm = -math.log(n + m)
    log_a = torch.full((n + 1,), norm, dtype=dtype); log_a[-1] += math.log(m)
    log_b = torch.full((m + 1,), norm, dtype=dtype); log_b[-1] += math.log(n)
    v = torch.zeros(B, m + 1, dtype=dtype)
    u = torch.zeros(B, n + 1, dtype=dtype)
    for _ in range(cfg.sinkhorn_iters):
        t = z + v[:, None, :]
        mx = t.max(2, keepdim=True).values
        e = (t - mx).exp()
        S = e.sum(2, keepdim=True)
        u = log_a - (mx[..., 0] + S[..., 0].log())
        c = (e * (log_a.exp()[None, :, None] / S)).sum(1)
        v = log_b + v - c.log()
    scores = z + u[:, :, None] + v[:, None, :] - norm
    return {'scores': scores, 'context_descriptors0': g0.transpose(1, 2), 'context_descriptors1': g1.transpose(1, 2)}
