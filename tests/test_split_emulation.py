"""CPU check of the operand split the tensor-core kernels use (csrc/tc_common.cuh: split_tf32_fast): hi = x rounded to tf32,
lo = x - hi handed to the tensor core unrounded, which then keeps only lo's top 10 mantissa bits (truncation).  The 3xTF32
product lo.hi + hi.lo + hi.hi built that way must stay far below the error of a plain fp32 GEMM and within a small factor of
the fully rounded split."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import superglue_oracle as O          # noqa: E402  (checker only)


def _truncate(x, bits=10):
    i = x.contiguous().view(torch.int32)
    return (i & ~((1 << (23 - bits)) - 1)).view(torch.float32)


def _x3(a, b, lo_of):
    ah, bh = O._round_mantissa(a, 10), O._round_mantissa(b, 10)
    al, bl = lo_of(a - ah), lo_of(b - bh)
    return al.double() @ bh.double() + ah.double() @ bl.double() + ah.double() @ bh.double()


def test_truncated_lo_split_is_fp32_grade():
    g = torch.Generator().manual_seed(0)
    for k in (64, 512, 2048):
        a, b = torch.randn(192, k, generator=g) * 3, torch.randn(k, 160, generator=g)
        ref = a.double() @ b.double()
        scale = ref.abs().max()
        err_rna = ((_x3(a, b, lambda r: O._round_mantissa(r, 10)) - ref).abs().max() / scale).item()
        err_trunc = ((_x3(a, b, _truncate) - ref).abs().max() / scale).item()
        err_fp32 = (((a @ b).double() - ref).abs().max() / scale).item()
        assert err_trunc < 2.0 * err_rna + 1e-9, (k, err_trunc, err_rna)
        assert err_trunc < 0.5 * err_fp32, (k, err_trunc, err_fp32)
        # positive operands (the softmax probabilities of P.V): truncation always shrinks |lo| - the bias must stay small too
        p = torch.rand(192, k, generator=g)
        refp = p.double() @ b.double().abs()
        errp = ((_x3(p, b.abs(), _truncate) - refp).abs().max() / refp.abs().max()).item()
        assert errp < 4e-7, (k, errp)
