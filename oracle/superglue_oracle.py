"""CPU oracle for the OpenGlue matching core  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, as plain functions over a ``state_dict``, the algorithm of the
reference's hot path (ucuapps/OpenGlue @ de2a26a).  It is the checker that the CUDA
path in ``openglue_b200/`` is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it; the product package never does
(tests/test_host_logic.py::test_product_never_imports_the_oracle greps for that).

Parity pin: the reference ships NO golden vectors or tests, so
the oracle is pinned against outputs of the reference itself, produced in the
build container by ``oracle/gen_golden.py`` (which imports the unmodified
reference from /root/reference) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every fixture.

Every function cites the reference lines it follows.  The arithmetic deliberately
uses the same ATen primitives as the reference (conv1d k=1, matmul, softmax,
logsumexp) and the same channel-first [B, C, n] activation layout, so that in
fp32 it reproduces the reference to rounding noise and, timed on host cores, is
a faithful stand-in for "the reference's PyTorch-CPU path".

An optional ``Contraction`` hook lets tests emulate reduced-precision tensor-core
operand rounding (tf32 / 3xtf32 / bf16 ...) in every GEMM-shaped contraction, to
predict what a given kernel precision does to the final log-scores.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# contraction hook (precision emulation)
# --------------------------------------------------------------------------- #
class Contraction:
    """All GEMM-shaped work in the oracle goes through these two methods.

    The default implementation is exactly what the reference executes
    (``F.conv1d`` with kernel 1 and ``torch.matmul``)."""

    def conv1x1(self, x: Tensor, w: Tensor, b: Optional[Tensor], tag: str) -> Tensor:
        return F.conv1d(x, w, b)

    def matmul(self, a: Tensor, b: Tensor, tag: str) -> Tensor:
        return torch.matmul(a, b)


def _round_mantissa(x: Tensor, keep_bits: int) -> Tensor:
    """Round-to-nearest-even an fp32 tensor to ``keep_bits`` explicit mantissa bits."""
    if x.dtype != torch.float32:
        return x
    drop = 23 - keep_bits
    xi = x.contiguous().view(torch.int32)
    bias = ((xi >> drop) & 1) + ((1 << (drop - 1)) - 1)
    yi = (xi + bias) & ~((1 << drop) - 1)
    return yi.view(torch.float32)


class EmulatedContraction(Contraction):
    """Operand-rounding emulation of tensor-core arithmetic (fp32 accumulate).

    mode per tag-prefix: 'fp32' | 'tf32' | 'tf32x3' | 'bf16' | 'bf16x3' | 'fp16x3' | 'fp16x3t'.
    'fp16x3': both operands scaled by a per-tensor power of two so that a BOUND on max|x| lands in [2^14, 2^15)
    (the bound is ``2**headroom`` times the true maximum: the kernels derive it from norms, not from the data),
    split into hi = fp16(x), lo = fp16(x - hi) (IEEE half incl. subnormals: absolute floor 2^-25 after scaling),
    three products with fp32 accumulation, un-scaled afterwards.  'fp16x3t': lo is TRUNCATED to fp16 instead of
    rounded (what a cvt.rz / bit-mask split would do).
    ``modes`` maps a tag prefix ('proj', 'qk', 'pv', 'mlp', 'final', 'score',
    'kenc') to a mode; '*' is the default."""

    def __init__(self, modes: Dict[str, str], headroom: int = 6):
        self.modes = modes
        self.headroom = headroom

    def _split_f16(self, x: Tensor, truncate_lo: bool = False):
        amax = float(x.abs().max())
        if amax == 0.0 or not math.isfinite(amax):
            return x, torch.zeros_like(x), 1.0
        e = 14 - self.headroom - math.floor(math.log2(amax))        # amax * 2^e in [2^(14-h), 2^(15-h))
        sc = 2.0 ** e
        xs = x * sc
        hi = xs.half().float() if x.dtype == torch.float32 else xs.half().to(x.dtype)
        r = xs - hi
        if truncate_lo:
            lo = r.half()
            over = lo.to(r.dtype).abs() > r.abs()
            lo = torch.where(over, torch.nextafter(lo, torch.zeros_like(lo)), lo).to(x.dtype)
        else:
            lo = r.half().to(x.dtype)
        return hi, lo, sc

    def _mode(self, tag: str) -> str:
        for k, v in self.modes.items():
            if k != '*' and tag.startswith(k):
                return v
        return self.modes.get('*', 'fp32')

    @staticmethod
    def _split(x: Tensor, bits: int, terms: int):
        parts, r = [], x
        for _ in range(terms):
            h = _round_mantissa(r, bits)
            parts.append(h)
            r = r - h
        return parts

    def _mm(self, a: Tensor, b: Tensor, mode: str) -> Tensor:
        if mode == 'fp32':
            return torch.matmul(a, b)
        if mode in ('fp16x3', 'fp16x3t'):
            ah, al, sa = self._split_f16(a, mode == 'fp16x3t')
            bh, bl, sb = self._split_f16(b, mode == 'fp16x3t')
            return (torch.matmul(al, bh) + torch.matmul(ah, bl) + torch.matmul(ah, bh)) * (1.0 / (sa * sb))
        bits = 10 if mode.startswith('tf32') else 7
        if mode in ('tf32', 'bf16'):
            return torch.matmul(_round_mantissa(a, bits), _round_mantissa(b, bits))
        if mode in ('tf32x3', 'bf16x3'):
            (ah, al), (bh, bl) = self._split(a, bits, 2), self._split(b, bits, 2)
            return torch.matmul(al, bh) + torch.matmul(ah, bl) + torch.matmul(ah, bh)
        if mode in ('tf32x2a', 'bf16x2a'):      # only the A operand is split
            (ah, al), bh = self._split(a, bits, 2), _round_mantissa(b, bits)
            return torch.matmul(al, bh) + torch.matmul(ah, bh)
        if mode in ('tf32x2b', 'bf16x2b'):      # only the B operand is split
            ah, (bh, bl) = _round_mantissa(a, bits), self._split(b, bits, 2)
            return torch.matmul(ah, bl) + torch.matmul(ah, bh)
        raise ValueError(mode)

    def conv1x1(self, x, w, b, tag):
        y = self._mm(w[:, :, 0], x, self._mode(tag))
        return y if b is None else y + b[None, :, None]

    def matmul(self, a, b, tag):
        return self._mm(a, b, self._mode(tag))


_DEFAULT = Contraction()


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def normalize_keypoints(kpts: Tensor, height: int, width: int) -> Tensor:
    """reference models/superglue/superglue.py:74-78 :  2*k / [W-1, H-1] - 1."""
    size = torch.tensor([width - 1, height - 1], device=kpts.device)
    return 2 * kpts / size - 1.0


def feed_forward_net(x: Tensor, sd: StateDict, prefix: str, n_linear: int,
                     mm: Contraction = _DEFAULT, tag: str = 'mlp', eps: float = 1e-5) -> Tensor:
    """reference models/utils.py:48-58 (FeedForwardNet): for every hidden layer
    Conv1d(k=1) -> ReLU -> BatchNorm1d (ReLU BEFORE BN), then a final Conv1d.
    Sequential indices: conv 3i, relu 3i+1, bn 3i+2.  Eval-mode BN (running stats)."""
    for i in range(n_linear - 1):
        x = mm.conv1x1(x, sd[f'{prefix}{3 * i}.weight'], sd[f'{prefix}{3 * i}.bias'], f'{tag}.{i}')
        x = torch.relu(x)
        bn = f'{prefix}{3 * i + 2}.'
        x = F.batch_norm(x, sd[bn + 'running_mean'], sd[bn + 'running_var'],
                         sd[bn + 'weight'], sd[bn + 'bias'], training=False, eps=eps)
    j = 3 * (n_linear - 1)
    return mm.conv1x1(x, sd[f'{prefix}{j}.weight'], sd[f'{prefix}{j}.bias'], f'{tag}.{n_linear - 1}')


def keypoint_encoder(kpts_n: Tensor, side_info: Tensor, sd: StateDict, n_linear: int,
                     mm: Contraction = _DEFAULT) -> Tensor:
    """reference models/superglue/positional_encoding.py:16-19:
    cat([kpts, side_info], -1) -> [B, 2+S, n] -> FeedForwardNet."""
    x = torch.cat([kpts_n, side_info], dim=-1).transpose(1, 2).contiguous()
    return feed_forward_net(x, sd, 'positional_encoding.encoder.', n_linear, mm, tag='kenc')


def softmax_attention(q: Tensor, k: Tensor, v: Tensor, mm: Contraction = _DEFAULT) -> Tensor:
    """reference models/superglue/attention.py:8-19.  q [B,H,Dh,N], k,v [B,H,Dh,M].
    softmax over the source keypoints (last dim of QK^T); returns [B,H,Dh,N].
    (The reference also returns the probabilities; its only caller drops them.)"""
    dh = q.size(2)
    att = mm.matmul(q.transpose(2, 3).contiguous(), k, 'qk') * dh ** -0.5
    att = att.softmax(dim=-1)
    out = mm.matmul(att, v.transpose(2, 3).contiguous(), 'pv')
    return out.transpose(2, 3).contiguous()


def multihead_attention(xq: Tensor, xs: Tensor, sd: StateDict, prefix: str, num_heads: int,
                        mm: Contraction = _DEFAULT) -> Tensor:
    """reference models/superglue/attention_gnn.py:22-32.  Heads are CONTIGUOUS channel
    blocks: view(B, H, Dh, n) => channel c = h*Dh + j."""
    b, d, _ = xq.shape
    dh = d // num_heads
    q = mm.conv1x1(xq, sd[prefix + 'in_proj_q.weight'], sd[prefix + 'in_proj_q.bias'], 'proj.q')
    k = mm.conv1x1(xs, sd[prefix + 'in_proj_k.weight'], sd[prefix + 'in_proj_k.bias'], 'proj.k')
    v = mm.conv1x1(xs, sd[prefix + 'in_proj_v.weight'], sd[prefix + 'in_proj_v.bias'], 'proj.v')
    o = softmax_attention(q.view(b, num_heads, dh, -1), k.view(b, num_heads, dh, -1),
                          v.view(b, num_heads, dh, -1), mm)
    o = o.view(b, d, -1)
    return mm.conv1x1(o, sd[prefix + 'out_proj.weight'], sd[prefix + 'out_proj.bias'], 'proj.o')


def propagate(xq: Tensor, xs: Tensor, sd: StateDict, prefix: str, num_heads: int,
              use_offset: bool, mm: Contraction = _DEFAULT) -> Tensor:
    """reference attention_gnn.py:43-55:  xq + fc(cat[xq, mha(xq, xs, xs)])
    (cat[xq - msg, msg] with use_offset); fc = FeedForwardNet(2d, 2d, d)."""
    msg = multihead_attention(xq, xs, sd, prefix + 'mha.', num_heads, mm)
    cat = torch.cat([xq - msg if use_offset else xq, msg], dim=1)
    return xq + feed_forward_net(cat, sd, prefix + 'fc.', 2, mm, tag='mlp')


def attention_gnn(x0: Tensor, x1: Tensor, sd: StateDict, num_stages: int, num_heads: int,
                  use_offset: bool, mm: Contraction = _DEFAULT) -> Tuple[Tensor, Tensor]:
    """reference attention_gnn.py:58-93.  Layer 2s = self, 2s+1 = cross.  The cross
    update is SEQUENTIAL (attention_gnn.py:74-77): image 1 attends to the already
    updated image 0."""
    for layer in range(2 * num_stages):
        p = f'attention_gnn.layers.{layer}.module.'
        if layer % 2 == 0:
            x0 = propagate(x0, x0, sd, p, num_heads, use_offset, mm)
            x1 = propagate(x1, x1, sd, p, num_heads, use_offset, mm)
        else:
            x0 = propagate(x0, x1, sd, p, num_heads, use_offset, mm)
            x1 = propagate(x1, x0, sd, p, num_heads, use_offset, mm)
    return x0, x1


def sinkhorn_log(log_a: Tensor, log_b: Tensor, z: Tensor, num_iters: int, reg: float) -> Tensor:
    """reference models/superglue/optimal_transport.py:4-28 (log_otp_solver):
    Z /= reg; u = v = 0; T x { u = log_a - LSE_j(Z + v);  v = log_b - LSE_i(Z + u) };
    returns Z + u + v.  (u first; v sees the new u.)"""
    z = z / reg
    u, v = torch.zeros_like(log_a), torch.zeros_like(log_b)
    for _ in range(num_iters):
        u = log_a - torch.logsumexp(z + v.unsqueeze(1), dim=2)
        v = log_b - torch.logsumexp(z + u.unsqueeze(2), dim=1)
    return z + u.unsqueeze(2) + v.unsqueeze(1)


def matching_log_probs(s: Tensor, dustbin: Tensor, num_iters: int, reg: float) -> Tensor:
    """reference superglue.py:88-111 (get_matching_probs).  s [B, m, n]  (m = #kpts
    of image 0).  Dustbin row/column = dustbin score; log_a = -log(m+n) with the
    last entry + log(n); log_b likewise + log(m); result = sinkhorn - norm."""
    bsz, m, n = s.shape
    z = torch.empty(bsz, m + 1, n + 1, dtype=s.dtype, device=s.device)
    z[:, :m, :n] = s
    z[:, m, :] = dustbin
    z[:, :, n] = dustbin
    # NB: like the reference, `norm`, log_a and log_b are float32 whatever s.dtype is
    # (torch.tensor(int).log() -> float32); they promote when combined with fp64 scores.
    norm = -torch.tensor(n + m, device=s.device).log()
    log_a = norm.expand(m + 1).contiguous()
    log_b = norm.expand(n + 1).contiguous()
    log_a[-1] += math.log(n)
    log_b[-1] += math.log(m)
    log_p = sinkhorn_log(log_a.expand(bsz, -1), log_b.expand(bsz, -1), z, num_iters, reg)
    return log_p - norm


# --------------------------------------------------------------------------- #
# the full path
# --------------------------------------------------------------------------- #
def image_hw(data: dict, idx: int) -> Tuple[int, int]:
    """reference superglue.py:35-38: image tensor [..., H, W] or image{idx}_size = (W, H)."""
    if 'image0' in data and 'image1' in data:
        sz = data[f'image{idx}'].size()
        return int(sz[-2]), int(sz[-1])
    w, h = data[f'image{idx}_size'][:2]
    return int(h), int(w)


def superglue_forward(sd: StateDict, config: dict, data: dict,
                      mm: Contraction = _DEFAULT) -> Dict[str, Tensor]:
    """reference superglue.py:29-72 (SuperGlue.forward), eval mode."""
    d = config['descriptor_dim']
    pe_cfg, gnn_cfg = config['positional_encoding'], config['attention_gnn']
    n_linear = len(pe_cfg.get('hidden_layers_sizes') or []) + 1
    ldesc0 = data['local_descriptors0'].transpose(1, 2).contiguous()
    ldesc1 = data['local_descriptors1'].transpose(1, 2).contiguous()
    h0, w0 = image_hw(data, 0)
    h1, w1 = image_hw(data, 1)
    k0 = normalize_keypoints(data['keypoints0'], h0, w0)
    k1 = normalize_keypoints(data['keypoints1'], h1, w1)
    pe0 = keypoint_encoder(k0, data['side_info0'], sd, n_linear, mm)
    pe1 = keypoint_encoder(k1, data['side_info1'], sd, n_linear, mm)
    if config.get('no_descriptors', False):
        x0, x1 = pe0, pe1
    else:
        x0, x1 = ldesc0 + pe0, ldesc1 + pe1
    x0, x1 = attention_gnn(x0, x1, sd, gnn_cfg['num_stages'], gnn_cfg['num_heads'],
                           gnn_cfg.get('use_offset', False), mm)
    g0 = mm.conv1x1(x0, sd['linear_proj.weight'], sd['linear_proj.bias'], 'final')
    g1 = mm.conv1x1(x1, sd['linear_proj.weight'], sd['linear_proj.bias'], 'final')
    if config.get('residual', False):
        alpha = torch.sigmoid(sd['mix_coefs'])
        g0 = alpha * g0 + (1.0 - alpha) * ldesc0
        g1 = alpha * g1 + (1.0 - alpha) * ldesc1
    s = mm.matmul(g0.transpose(1, 2).contiguous(), g1, 'score') * d ** -0.5
    scores = matching_log_probs(s, sd['dustbin_score'], config['otp']['num_iters'], config['otp']['reg'])
    return {'context_descriptors0': g0, 'context_descriptors1': g1, 'scores': scores}


def extract_matches(scores: Tensor, match_threshold: float) -> Dict[str, Tensor]:
    """reference models/matching_module.py:174-187 (+ the reverse direction of
    inference.py:176-190).  torch.max(dim) returns the FIRST maximal index."""
    inner = scores[:, :-1, :-1]
    max0, max1 = inner.max(2), inner.max(1)
    i0, i1 = max0.indices, max1.indices
    ar0 = torch.arange(i0.shape[1], device=i0.device)[None]
    ar1 = torch.arange(i1.shape[1], device=i1.device)[None]
    mutual0 = ar0 == i1.gather(1, i0)
    mutual1 = ar1 == i0.gather(1, i1)
    zero = scores.new_tensor(0)
    ms0 = torch.where(mutual0, max0.values.exp(), zero)
    ms1 = torch.where(mutual1, ms0.gather(1, i1), zero)
    valid0 = mutual0 & (ms0 > match_threshold)
    valid1 = mutual1 & valid0.gather(1, i1)
    return {
        'matches0': torch.where(valid0, i0, i0.new_tensor(-1)),
        'matching_scores0': ms0,
        'matches1': torch.where(valid1, i1, i1.new_tensor(-1)),
        'matching_scores1': ms1,
    }


def cast_state_dict(sd: StateDict, dtype: torch.dtype) -> StateDict:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def cast_data(data: dict, dtype: torch.dtype) -> dict:
    return {k: (v.to(dtype) if isinstance(v, Tensor) and v.is_floating_point() else v)
            for k, v in data.items()}


def run(sd: StateDict, config: dict, data: dict, match_threshold: float = 0.2,
        dtype: torch.dtype = torch.float32, mm: Contraction = _DEFAULT) -> Dict[str, Tensor]:
    """Whole hot path: SuperGlue.forward + match extraction, under no_grad."""
    with torch.no_grad():
        out = superglue_forward(cast_state_dict(sd, dtype), config, cast_data(data, dtype), mm)
        out.update(extract_matches(out['scores'], match_threshold))
    return out
