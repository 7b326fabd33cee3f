"""Training-mode forward AND backward of the matching core: what ``SuperGlue.forward`` and
torch autograd compute for the reference inside ``MatchingTrainingModule.training_step`` (reference
models/matching_module.py:71-105) - BatchNorm with batch statistics (models/utils.py:48-58: Conv1d -> ReLU -> BatchNorm1d),
running-statistics updates, and the gradient of every parameter (and of the local descriptors).

Nothing here is differentiated by torch: :class:`TrainStep` records the activations of the forward pass and runs the backward
pass as an explicit schedule of the library's kernels (``include/openglue_b200.h``: the tcgen05 3xTF32 / fp32 GEMMs for
``dX = dY W`` and ``dW = dY^T X``, the fused attention kernel forward, the materialised-softmax attention gradient, the
Sinkhorn backward pass, batch-norm / bias / mix reductions).  torch only owns the buffers and routes the resulting gradients
to the ``nn.Parameter`` objects (``_TrainFunction``).  There is no CPU path.

Layout: every activation is row-major ``[rows, channels]`` with ``rows = batch x keypoints`` of ONE image - the reference's
``[B, C, N]`` tensors transposed; BatchNorm statistics therefore run over the rows of one call, exactly the reference's
per-call ``(B, N)`` statistics (attention_gnn.py:58-77 calls the shared module once per image).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _cabi
from ._ops import _Ops, _p, _pad4  # noqa: F401  (tests build their torch double of the kernels on _Ops' composite helpers)

__all__ = ['TrainStep', 'train_forward', 'GraphedTrainStep']


class _BN:
    """One BatchNorm1d call site: parameters, running buffers and what the backward pass needs."""

    def __init__(self, mod: torch.nn.BatchNorm1d):
        self.mod = mod

    def forward(self, ops: _Ops, a: torch.Tensor) -> torch.Tensor:
        m = self.mod
        self.a = a
        if m.momentum is None:
            raise NotImplementedError('BatchNorm1d(momentum=None) (cumulative average) is not built')
        track = m.track_running_stats and m.running_mean is not None
        y, self.mean, self.invstd = ops.bn_fwd(a, m.weight, m.bias, m.eps, m.momentum, m.running_mean if track else None,
                                               m.running_var if track else None)
        if track:
            m.num_batches_tracked += 1
        return y

    def backward(self, ops: _Ops, dy: torch.Tensor):
        return ops.bn_bwd(dy, self.a, self.mod.weight, self.mean, self.invstd)


class TrainStep:
    """One training-mode forward pass of ``model`` (an ``openglue_b200.SuperGlue`` in ``train()`` mode) with everything its
    backward pass needs.  ``forward()`` -> (scores [B,N+1,M+1], ctx0 [B,d,N], ctx1 [B,d,M]); ``backward(dscores, dctx0, dctx1)`` ->
    ``{parameter name: gradient}`` (+ ``'local_descriptors0/1'``)."""

    def __init__(self, model, data: dict, ops=None):
        self.model = model
        cfg = model.config
        self.d = cfg['descriptor_dim']
        self.H = cfg['attention_gnn']['num_heads']
        self.use_offset = bool(cfg['attention_gnn'].get('use_offset', False))
        self.residual = bool(model.residual)
        self.no_desc = bool(cfg.get('no_descriptors', False))
        self.iters, self.reg = int(cfg['otp']['num_iters']), float(cfg['otp']['reg'])
        self.S = cfg['positional_encoding'].get('side_info_size', 1)
        k0 = data['keypoints0']
        self.dev = k0.device
        if ops is None:
            if self.dev.type != 'cuda':
                raise RuntimeError('openglue_b200 training needs CUDA tensors (sm_100a); there is no CPU path')
            for p in model.parameters():
                if p.device != self.dev or p.dtype != torch.float32:
                    raise RuntimeError('openglue_b200 training needs float32 parameters on the device of the data')
            prec = model._precision()
            ops = _Ops(self.dev, _cabi.OG_PREC_FP32 if prec == _cabi.OG_PREC_FP32 else _cabi.OG_PREC_TF32X3)
        self.ops = ops                                  # (tests inject a torch double of the kernels to check this schedule on the CPU)
        f = lambda t, last: self._prep(t, last)
        self.kpts = [f(data['keypoints0'], 2), f(data['keypoints1'], 2)]
        self.side = [f(data['side_info0'], self.S), f(data['side_info1'], self.S)]
        self.ldesc = [f(data['local_descriptors0'], self.d), f(data['local_descriptors1'], self.d)]
        self.B = self.kpts[0].shape[0]
        self.N = [self.kpts[0].shape[1], self.kpts[1].shape[1]]
        if min(self.N) == 0:
            raise ValueError('empty keypoint set')
        self.wh = [model._image_wh(data, 0), model._image_wh(data, 1)]
        self.grads: Dict[str, torch.Tensor] = {}

    @staticmethod
    def _prep(t, last):
        t = t.detach()
        if t.dtype != torch.float32:
            t = t.float()
        if t.shape[-1] != last:
            raise ValueError(f'expected last dimension {last}, got {tuple(t.shape)}')
        return t.contiguous()

    # ------------------------------------------------------------------ helpers
    def _device_ctx(self):
        import contextlib
        return torch.cuda.device(self.dev) if self.dev.type == 'cuda' else contextlib.nullcontext()

    @staticmethod
    def _w2(conv):
        return conv.weight.view(conv.weight.shape[0], conv.weight.shape[1])

    def _acc(self, name: str, g: torch.Tensor):
        """grads[name] += g (parameters shared by both images / both cross directions)"""
        if name in self.grads:
            self.ops.axpby(self.grads[name], g, 1.0, 1.0, out=self.grads[name])
        else:
            self.grads[name] = g

    def _acc_weight(self, name: str, conv, dY, X, X2=None):
        """d conv.weight (+)= dY^T [X | X2], d conv.bias (+)= colsum dY"""
        ops = self.ops
        wname, bname = name + '.weight', name + '.bias'
        if wname not in self.grads:
            self.grads[wname] = ops.zeros(conv.weight.shape[0], conv.weight.shape[1])
        ops.grad_weight(dY, X, self.grads[wname])
        if X2 is not None:
            ops.grad_weight(dY, X2, self.grads[wname], col_off=X.shape[1])
        self._acc(bname, ops.colsum(dY))

    # ------------------------------------------------------------------ forward
    def forward(self):
        ops, model, B, d = self.ops, self.model, self.B, self.d
        with self._device_ctx():
            enc = model.positional_encoding.encoder
            nl = (len(enc) + 2) // 3                                   # Conv (ReLU BN Conv)*
            self.kenc = []                                             # per image: list of (input, _BN) per hidden layer + last input
            x = []
            for i in range(2):
                rows = B * self.N[i]
                h = ops.kenc_input(self.kpts[i], self.side[i], rows, self.S, self.wh[i][0], self.wh[i][1])
                rec = []
                for j in range(nl - 1):
                    conv, bn = enc[3 * j], _BN(enc[3 * j + 2])
                    a = ops.linear(h, self._w2(conv), conv.bias)
                    rec.append((h, bn))
                    h = bn.forward(ops, a)
                conv = enc[3 * (nl - 1)]
                xi = ops.linear(h, self._w2(conv), conv.bias, R=None if self.no_desc else self.ldesc[i].view(rows, d))
                rec.append((h, None))
                self.kenc.append(rec)
                x.append(xi)
            self.calls = []                                            # message-passing calls in execution order
            for l, layer in enumerate(model.attention_gnn.layers):
                mod = layer.module
                name = f'attention_gnn.layers.{l}.module'
                if l % 2 == 0:                                         # self (attention_gnn.py:58-61)
                    x[0] = self._prop(name, mod, x[0], x[0], 0, 0)
                    x[1] = self._prop(name, mod, x[1], x[1], 1, 1)
                else:                                                  # cross, sequential (attention_gnn.py:74-77)
                    x[0] = self._prop(name, mod, x[0], x[1], 0, 1)
                    x[1] = self._prop(name, mod, x[1], x[0], 1, 0)
            self.x_final = x
            proj = model.linear_proj
            self.g = [ops.linear(x[i], self._w2(proj), proj.bias) for i in range(2)]
            if self.residual:
                self.mix = model.mix_coefs.detach().reshape(d)
                self.m = [ops.mix_fwd(self.g[i], self.ldesc[i].view(B * self.N[i], d), self.mix) for i in range(2)]
            else:
                self.m = self.g
            n, m_ = self.N
            # context descriptors in the reference's [B, d, N] layout; their zero-padded copies are the transposed operands of the backward pass
            self.mT = [ops.transpose(self.m[i], batch=B, rows=self.N[i], cols=d) for i in range(2)]          # [B, d, pad4(N)]
            ctx = [self.mT[i][:, :, :self.N[i]].contiguous() if self.mT[i].shape[2] != self.N[i] else self.mT[i] for i in range(2)]
            lds = _pad4(m_)
            self.Sp = ops.zeros(B, n, lds)
            ops.gemm(self.m[0], d, d, self.m[1], d, n, m_, self.Sp, lds, alpha=d ** -0.5, batch=B, strideA=n * d, strideW=m_ * d, strideY=n * lds)
            self.dust = model.dustbin_score.detach().reshape(1).contiguous()
            scores, self.hist = ops.sinkhorn_fwd(self.Sp, self.dust, B, n, m_, self.iters, self.reg)
        return scores, ctx[0], ctx[1]

    def _prop(self, name, mod, xq, xkv, iq, ikv):
        """ResidualAttentionMessagePropagation.forward (attention_gnn.py:43-55) on row-major activations."""
        ops, d, H, B = self.ops, self.d, self.H, self.B
        nq, nk = self.N[iq], self.N[ikv]
        mha = mod.mha
        q = ops.linear(xq, self._w2(mha.in_proj_q), mha.in_proj_q.bias)
        k = ops.linear(xkv, self._w2(mha.in_proj_k), mha.in_proj_k.bias)
        v = ops.linear(xkv, self._w2(mha.in_proj_v), mha.in_proj_v.bias)
        o = ops.attention(q, k, v, B, nq, nk, H, d // H)
        msg = ops.linear(o, self._w2(mha.out_proj), mha.out_proj.bias)
        c1 = ops.axpby(xq, msg, 1.0, -1.0) if self.use_offset else xq
        fc = mod.fc
        a = ops.linear(c1, self._w2(fc[0]), fc[0].bias, A2=msg)
        bn = _BN(fc[2])
        hbn = bn.forward(ops, a)
        out = ops.linear(hbn, self._w2(fc[3]), fc[3].bias, R=xq)
        self.calls.append(dict(name=name, mod=mod, xq=xq, xkv=xkv, iq=iq, ikv=ikv, q=q, k=k, v=v, o=o, msg=msg, c1=c1, bn=bn, hbn=hbn))
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, dscores: Optional[torch.Tensor], dctx0: Optional[torch.Tensor] = None, dctx1: Optional[torch.Tensor] = None
                 ) -> Dict[str, torch.Tensor]:
        ops, model, B, d = self.ops, self.model, self.B, self.d
        n, m_ = self.N
        self.grads = {}
        with self._device_ctx():
            dm = [ops.zeros(B * n, d), ops.zeros(B * m_, d)]
            if dscores is not None:
                G = dscores.detach().float().contiguous()
                dZ, dd = ops.sinkhorn_bwd(self.Sp, self.dust, self.hist, G, B, n, m_, self.iters, self.reg)
                self.grads['dustbin_score'] = dd.reshape(model.dustbin_score.shape)
                # S = m0 m1^T d^-0.5:  dm0 = dS m1 d^-0.5,  dm1 = dS^T m0 d^-0.5   (superglue.py:64, 80-85)
                mp, np_ = _pad4(m_), _pad4(n)
                dS = ops.zeros(B, n, mp)
                ops.transpose_raw(dZ, 0, m_ + 1, (n + 1) * (m_ + 1), dS, mp, n * mp, B, n, m_, False)
                dSt = ops.zeros(B, m_, np_)
                ops.transpose_raw(dZ, 0, m_ + 1, (n + 1) * (m_ + 1), dSt, np_, m_ * np_, B, n, m_, True)
                sc = d ** -0.5
                ops.gemm(dS, mp, mp, self.mT[1], mp, n, d, dm[0], d, alpha=sc, batch=B, strideA=n * mp, strideW=d * mp, strideY=n * d)
                ops.gemm(dSt, np_, np_, self.mT[0], np_, m_, d, dm[1], d, alpha=sc, batch=B, strideA=m_ * np_, strideW=d * np_, strideY=m_ * d)
            for i, dctx in enumerate((dctx0, dctx1)):
                if dctx is not None:                                    # gradient arriving at the [B, d, N] context descriptors
                    t = ops.transpose(dctx.detach().float().contiguous(), batch=B, rows=d, cols=self.N[i], pad=False)     # -> [B, N, d]
                    ops.axpby(dm[i], t.view(B * self.N[i], d), 1.0, 1.0, out=dm[i])
            dl = [None, None]
            if self.residual:
                dg = []
                csum = None
                for i in range(2):
                    rows = B * self.N[i]
                    gi, li = ops.mix_bwd(dm[i], self.mix)
                    dg.append(gi)
                    dl[i] = li
                    c = ops.colsum(dm[i], self.g[i], self.ldesc[i].view(rows, d))
                    csum = c if csum is None else ops.axpby(csum, c, 1.0, 1.0)
                self.grads['mix_coefs'] = ops.mix_param_grad(csum, self.mix).reshape(model.mix_coefs.shape)
            else:
                dg = dm
            proj = model.linear_proj
            dx = []
            for i in range(2):
                self._acc_weight('linear_proj', proj, dg[i], self.x_final[i])
                dx.append(ops.grad_input(dg[i], self._w2(proj)))
            # message passing, in reverse execution order
            for call in reversed(self.calls):
                iq, ikv = call['iq'], call['ikv']
                dxq, dxkv = self._prop_bwd(call, dx[iq])
                if iq == ikv:
                    dx[iq] = ops.axpby(dxq, dxkv, 1.0, 1.0, out=dxq)
                else:
                    dx[iq] = dxq
                    dx[ikv] = ops.axpby(dx[ikv], dxkv, 1.0, 1.0, out=dxkv)
            # keypoint encoder (+ the descriptors that were added to its output, superglue.py:52-55)
            enc = model.positional_encoding.encoder
            for i in range(2):
                rec = self.kenc[i]
                if not self.no_desc:
                    dl[i] = dx[i] if dl[i] is None else ops.axpby(dl[i], dx[i], 1.0, 1.0, out=dl[i])
                g = dx[i]
                for j in range(len(rec) - 1, -1, -1):
                    h, bn = rec[j]
                    conv = enc[3 * j]
                    self._acc_weight(f'positional_encoding.encoder.{3 * j}', conv, g, h)
                    if j == 0:
                        break                                           # no gradient with respect to the keypoints / side info
                    gh = ops.grad_input(g, self._w2(conv))
                    g, dgam, dbet = rec[j - 1][1].backward(ops, gh)
                    self._acc(f'positional_encoding.encoder.{3 * (j - 1) + 2}.weight', dgam)
                    self._acc(f'positional_encoding.encoder.{3 * (j - 1) + 2}.bias', dbet)
            for i in range(2):
                if dl[i] is not None:
                    self.grads[f'local_descriptors{i}'] = dl[i].view(B, self.N[i], d)
        return self.grads

    def _prop_bwd(self, call, dout):
        ops, d = self.ops, self.d
        mod, name = call['mod'], call['name']
        fc, mha = mod.fc, mod.mha
        xq, xkv = call['xq'], call['xkv']
        # out = xq + hbn W2^T + b2
        self._acc_weight(name + '.fc.3', fc[3], dout, call['hbn'])
        dhbn = ops.grad_input(dout, self._w2(fc[3]))
        da, dgam, dbet = call['bn'].backward(ops, dhbn)
        self._acc(name + '.fc.2.weight', dgam)
        self._acc(name + '.fc.2.bias', dbet)
        # a = [c1 | msg] W1^T + b1
        self._acc_weight(name + '.fc.0', fc[0], da, call['c1'], call['msg'])
        W1 = self._w2(fc[0])
        dc1 = ops.grad_input(da, W1, 0, d)
        dmsg = ops.grad_input(da, W1, d, d)
        if self.use_offset:                                             # c1 = xq - msg
            dmsg = ops.axpby(dmsg, dc1, 1.0, -1.0, out=dmsg)
        dxq = ops.axpby(dout, dc1, 1.0, 1.0, out=dc1)
        # msg = o Wo^T + bo
        self._acc_weight(name + '.mha.out_proj', mha.out_proj, dmsg, call['o'])
        do = ops.grad_input(dmsg, self._w2(mha.out_proj))
        dq, dk, dv = self._attention_bwd(call, do)
        self._acc_weight(name + '.mha.in_proj_q', mha.in_proj_q, dq, xq)
        self._acc_weight(name + '.mha.in_proj_k', mha.in_proj_k, dk, xkv)
        self._acc_weight(name + '.mha.in_proj_v', mha.in_proj_v, dv, xkv)
        ops.axpby(dxq, ops.grad_input(dq, self._w2(mha.in_proj_q)), 1.0, 1.0, out=dxq)
        dxkv = ops.grad_input(dk, self._w2(mha.in_proj_k))
        ops.axpby(dxkv, ops.grad_input(dv, self._w2(mha.in_proj_v)), 1.0, 1.0, out=dxkv)
        return dxq, dxkv

    def _attention_bwd(self, call, do):
        """Gradient of softmax_attention (models/superglue/attention.py:8-19) per head, with the probabilities re-materialised:
        P = softmax(q k^T s);  dV = P^T dO;  dP = dO V^T;  dS = s P (dP - rowsum(P dP));  dQ = dS K;  dK = dS^T Q."""
        ops, d, H, B = self.ops, self.d, self.H, self.B
        dh = d // H
        nq, nk = self.N[call['iq']], self.N[call['ikv']]
        q, k, v = call['q'], call['k'], call['v']
        nqp, nkp = _pad4(nq), _pad4(nk)
        scale = dh ** -0.5
        qT = ops.transpose(q, batch=B, rows=nq, cols=d)                 # [B, d, nqp]
        kT = ops.transpose(k, batch=B, rows=nk, cols=d)                 # [B, d, nkp]
        doT = ops.transpose(do, batch=B, rows=nq, cols=d)               # [B, d, nqp]
        dq, dk, dv = ops.empty(B * nq, d), ops.empty(B * nk, d), ops.empty(B * nk, d)
        P, dP = ops.zeros(B, nq, nkp), ops.zeros(B, nq, nkp)
        PT, dST = ops.zeros(B, nk, nqp), ops.zeros(B, nk, nqp)
        for h in range(H):
            c = h * dh
            ops.gemm(q, d, dh, k, d, nq, nk, P, nkp, a_off=c, w_off=c, alpha=scale, batch=B, strideA=nq * d, strideW=nk * d, strideY=nq * nkp)
            ops.softmax_rows(P, nkp, B * nq, nk)
            ops.gemm(do, d, dh, v, d, nq, nk, dP, nkp, a_off=c, w_off=c, batch=B, strideA=nq * d, strideW=nk * d, strideY=nq * nkp)
            ops.transpose_raw(P, 0, nkp, nq * nkp, PT, nqp, nk * nqp, B, nq, nk, True)
            ops.gemm(PT, nqp, nqp, doT, nqp, nk, dh, dv, d, w_off=c * nqp, y_off=c, batch=B, strideA=nk * nqp, strideW=d * nqp, strideY=nk * d)
            ops.softmax_bwd_rows(P, dP, nkp, B * nq, nk, scale)
            ops.gemm(dP, nkp, nkp, kT, nkp, nq, dh, dq, d, w_off=c * nkp, y_off=c, batch=B, strideA=nq * nkp, strideW=d * nkp, strideY=nq * d)
            ops.transpose_raw(dP, 0, nkp, nq * nkp, dST, nqp, nk * nqp, B, nq, nk, True)
            ops.gemm(dST, nqp, nqp, qT, nqp, nk, dh, dk, d, w_off=c * nqp, y_off=c, batch=B, strideA=nk * nqp, strideW=d * nqp, strideY=nk * d)
        return dq, dk, dv


class _TrainFunction(torch.autograd.Function):
    """Routes the explicit backward pass of :class:`TrainStep` into torch's autograd graph: inputs are the local descriptors and
    every parameter of the model, outputs the three tensors of the reference's ``forward``."""

    @staticmethod
    def forward(ctx, model, data, names, ld0, ld1, *params):
        ctx.set_materialize_grads(False)
        step = TrainStep(model, data)
        scores, c0, c1 = step.forward()
        ctx.step, ctx.names = step, names
        ctx.shapes = [p.shape for p in params]
        ctx.ld_dtypes = (ld0.dtype, ld1.dtype)
        return scores, c0, c1

    @staticmethod
    def backward(ctx, dscores, dc0, dc1):
        g = ctx.step.backward(dscores, dc0, dc1)
        ctx.step = None                                                 # the activations are released with the step
        need = ctx.needs_input_grad
        gl = [g.get('local_descriptors0'), g.get('local_descriptors1')]
        out: List[Optional[torch.Tensor]] = [None, None, None]
        for i in range(2):
            out.append(gl[i].to(ctx.ld_dtypes[i]) if need[3 + i] and gl[i] is not None else None)
        for j, nme in enumerate(ctx.names):
            gj = g.get(nme) if need[5 + j] else None
            out.append(None if gj is None else gj.reshape(ctx.shapes[j]))
        return tuple(out)


def train_forward(model, data: dict) -> Dict[str, torch.Tensor]:
    """``SuperGlue.forward`` in training mode: same outputs as the reference module (superglue.py:66-70), differentiable."""
    names, params = zip(*[(n, p) for n, p in model.named_parameters()])
    scores, c0, c1 = _TrainFunction.apply(model, data, names, data['local_descriptors0'], data['local_descriptors1'], *params)
    return {'context_descriptors0': c0, 'context_descriptors1': c1, 'scores': scores}


class GraphedTrainStep:
    """The whole training step of one batch shape - train-mode forward, ``criterion``, backward, gradients into ``param.grad`` -
    captured ONCE into a CUDA graph and replayed: the eager step issues ~5000 kernel launches from Python and is launch-rate-bound,
    the replay costs the kernels' own time.  Labels (``generate_gt_matches``) and the optimiser stay outside::

        step = GraphedTrainStep(model, data, y_true)          # model.train(); captures on the first batch
        for data, y_true in loader:                           # same shapes
            loss = step(data, y_true)                         # {'loss', 'metric_loss'}; gradients are in p.grad
            optimizer.step()                                  # in-place updates keep the captured parameter addresses valid

    The graph reads the parameters and BatchNorm buffers in place (so optimiser steps and running statistics carry over) and the
    inputs from static copies.  Re-create the object when shapes change or parameters are re-allocated (``.to()``, ``load_state_dict``
    keeps storage and is fine).  Results are bit-identical to the eager step (same kernels, same order)."""

    _KEYS = ('keypoints0', 'keypoints1', 'side_info0', 'side_info1', 'local_descriptors0', 'local_descriptors1')

    def __init__(self, model, data: dict, y_true: dict, nll_weight: float = 1.0):
        from .losses import _run as criterion_run
        if not model.training:
            raise RuntimeError('GraphedTrainStep captures the training-mode step: call model.train() first')
        dev = data['keypoints0'].device
        if dev.type != 'cuda':
            raise RuntimeError('openglue_b200 training needs CUDA tensors (sm_100a); there is no CPU path')
        self.model, self.dev = model, dev
        self.static = dict(data)
        for k in self._KEYS:
            self.static[k] = data[k].detach().float().contiguous().clone()
        self.gt = {k: y_true[k].to(device=dev, dtype=torch.int64).contiguous().clone() for k in ('gt_matches0', 'gt_matches1')}
        self.params = list(model.named_parameters())
        for _, p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)

        def run():
            step = TrainStep(model, self.static)
            scores, _, _ = step.forward()
            loss, dscores = criterion_run(self.gt, {'scores': scores}, True, float(nll_weight))
            grads = step.backward(dscores)
            for name, p in self.params:
                p.grad.copy_(grads[name].reshape(p.shape))
            return loss, scores

        with torch.cuda.device(dev):
            saved = [b.clone() for b in model.buffers()]                # the warm-up run must not count as a training step
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                run()                                                    # builds function attributes, allocator pools
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            for b, s in zip(model.buffers(), saved):
                b.copy_(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.scores = run()

    def __call__(self, data: dict, y_true: dict) -> Dict[str, torch.Tensor]:
        for k in self._KEYS:
            if tuple(data[k].shape) != tuple(self.static[k].shape):
                raise ValueError(f'{k}: shape {tuple(data[k].shape)} differs from the captured {tuple(self.static[k].shape)}')
            self.static[k].copy_(data[k], non_blocking=True)
        for k in self.gt:
            self.gt[k].copy_(y_true[k], non_blocking=True)
        self.graph.replay()
        return {'loss': self.loss[0], 'metric_loss': self.loss[1]}
