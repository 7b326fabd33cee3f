"""Training step: forward in train() mode + the explicit backward pass, against fixtures minted
from the UNMODIFIED reference in train() mode with torch autograd (oracle/gen_golden_train.py)."""
import os

import pytest
import torch

from openglue_b200.synthetic import synthetic_state_dict

TRAIN_CASES = ['train_small', 'train_offset', 'train_ragged']


def _load(name):
    here = os.path.dirname(os.path.abspath(__file__))
    return torch.load(os.path.join(here, 'golden', name + '.pt'), weights_only=False)


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_training_fixture_contents(name):
    fx = _load(name)
    cfg = fx['config']
    sd = synthetic_state_dict(cfg, seed=fx['weights_seed'])
    assert set(fx['bn_buffers']) == {k for k in sd if 'running_' in k}
    params = {k for k in sd if 'running_' not in k and 'num_batches_tracked' not in k}
    assert set(fx['grads']) == params                                   # a gradient for EVERY parameter of the reference module
    for k in params:
        assert fx['grads'][k].shape == sd[k].shape
    assert fx['grads_f32_vs_f64_max_abs'] < 1e-5
    b, n, _ = fx['data']['keypoints0'].shape
    m = fx['data']['keypoints1'].shape[1]
    assert fx['scores_f64'].shape == (b, n + 1, m + 1)
    for k, v in fx['buffers_after'].items():                            # every shared module runs once per image: two updates per step
        if k.endswith('num_batches_tracked'):
            assert int(v) == int(sd[k]) + 2


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'tf32x3'])
@pytest.mark.parametrize('name', TRAIN_CASES)
def test_training_step_matches_reference(name, precision):
    from openglue_b200 import SuperGlue, criterion
    dev = torch.device('cuda:0')
    fx = _load(name)
    cfg = dict(fx['config'], precision=precision)
    sd = synthetic_state_dict(fx['config'], seed=fx['weights_seed'])
    sd.update(fx['bn_buffers'])
    model = SuperGlue(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in fx['data'].items()}
    data['local_descriptors0'] = data['local_descriptors0'].clone().requires_grad_(True)
    data['local_descriptors1'] = data['local_descriptors1'].clone().requires_grad_(True)
    y_true = {'gt_matches0': fx['gt_matches0'].to(dev), 'gt_matches1': fx['gt_matches1'].to(dev)}
    out = model(data)
    # log-scores reach |s| ~ 170 here (reg = 0.5): the bound is 1e-4 or a small multiple of the reference's OWN fp32-vs-fp64 distance
    bound = max(1e-4, 4 * float((fx['scores_f32'].double() - fx['scores_f64']).abs().max()))
    err = float((out['scores'].detach().cpu().double() - fx['scores_f64']).abs().max())
    print(f'{name} {precision}: max |scores - ref64| {err:.2e} (bound {bound:.2e})')
    assert err <= bound
    assert (out['context_descriptors0'].detach().cpu().double() - fx['context_descriptors0_f64']).abs().max() <= 1e-4
    loss = criterion(y_true, out, margin=None)
    assert abs(float(loss['loss'].detach()) - float(fx['loss_f64'])) <= 1e-4 * max(1.0, abs(float(fx['loss_f64'])))
    loss['loss'].backward()
    # gradients: within 1e-3 relative (per tensor, 2-norm) of the reference's autograd; elementwise within 1e-3 of the tensor's scale
    worst = ('', 0.0)
    for k, p in model.named_parameters():
        ref = fx['grads'][k]
        assert p.grad is not None, k
        g = p.grad.detach().cpu()
        assert g.shape == ref.shape, k
        scale = float(ref.abs().max())
        if scale < 1e-9:
            assert float(g.abs().max()) < 1e-6, k
            continue
        r = _rel(g, ref)
        worst = max(worst, (k, r), key=lambda t: t[1])
        assert r <= 1e-3, (k, r)
        assert float((g - ref).abs().max()) <= 1e-3 * scale, k
    print(f'{name} {precision}: worst relative gradient error {worst[1]:.2e} ({worst[0]})')
    for i in (0, 1):
        assert _rel(data[f'local_descriptors{i}'].grad.cpu(), fx[f'dlocal_descriptors{i}_f64']) <= 1e-3
    # BatchNorm running buffers moved exactly as nn.BatchNorm1d moves them
    for k, v in model.named_buffers():
        ref = fx['buffers_after'][k]
        if k.endswith('num_batches_tracked'):
            assert int(v) == int(ref), k
        else:
            assert (v.cpu() - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max())), k


@pytest.mark.gpu
def test_training_step_updates_with_an_optimizer():
    """A few optimiser steps on one synthetic batch reduce the loss (the drop-in runs inside an ordinary torch training loop)."""
    from openglue_b200 import SuperGlue, criterion
    dev = torch.device('cuda:0')
    fx = _load('train_small')
    model = SuperGlue(dict(fx['config'], precision='tf32x3'))
    sd = synthetic_state_dict(fx['config'], seed=fx['weights_seed'])
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in fx['data'].items()}
    y_true = {'gt_matches0': fx['gt_matches0'].to(dev), 'gt_matches1': fx['gt_matches1'].to(dev)}
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = criterion(y_true, model(data), margin=None)['loss']
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    model.eval()                                                         # the fused inference path picks the updated weights up
    with torch.no_grad():
        s = model(data)['scores']
    assert torch.isfinite(s).all()


# ---------------------------------------------------------------------------------------------------------------------
# The SCHEDULE of the training step (which operator on which buffer, with which offsets / strides / accumulation) checked on
# the CPU: the same openglue_b200.training.TrainStep driven by a torch double of the kernels (test infrastructure only).
class _CpuOps:
    def __init__(self, dtype=torch.float64):
        self.dt = dtype

    def empty(self, *shape):
        return torch.full(tuple(shape), float('nan'), dtype=self.dt)     # reading an unwritten element poisons the result

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=self.dt)

    @staticmethod
    def _view(t, off, rows, cols, ld):
        flat = t.detach().reshape(-1)                                     # (a view keeps its offset into the storage: as_strided's offset is absolute)
        return torch.as_strided(flat, (rows, cols), (ld, 1), flat.storage_offset() + off)

    def gemm(self, A, lda, k1, W, ldw, rows, nout, Y, ldy, *, a_off=0, w_off=0, y_off=0, A2=None, lda2=0, k2=0, a2_off=0, bias=None,
             relu=False, alpha=1.0, R=None, ldr=0, r_off=0, batch=1, strideA=0, strideA2=0, strideW=0, strideY=0, strideR=0,
             Yt=None, ldyt=0, strideYt=0, yt_off=0):
        assert Yt is None
        for b in range(batch):
            a = self._view(A, a_off + b * strideA, rows, k1, lda).to(self.dt)
            if A2 is not None:
                a = torch.cat([a, self._view(A2, a2_off + b * strideA2, rows, k2, lda2).to(self.dt)], 1)
            w = self._view(W, w_off + b * strideW, nout, k1 + k2, ldw).to(self.dt)
            y = alpha * (a @ w.t())
            if bias is not None:
                y = y + bias.detach().to(self.dt)
            if relu:
                y = y.clamp_min(0)
            if R is not None:
                y = y + self._view(R, r_off + b * strideR, rows, nout, ldr)
            self._view(Y, y_off + b * strideY, rows, nout, ldy).copy_(y)

    def transpose_raw(self, X, x_off, ld_in, stride_in, out, ld_out, stride_out, batch, rows, cols, transpose):
        for b in range(batch):
            src = self._view(X, x_off + b * stride_in, rows, cols, ld_in)
            if transpose:
                self._view(out, b * stride_out, cols, rows, ld_out).copy_(src.t())
            else:
                self._view(out, b * stride_out, rows, cols, ld_out).copy_(src)

    def kenc_input(self, kpts, side, rows, S, width, height):
        k = kpts.reshape(rows, 2).to(self.dt)
        xy = 2 * k / torch.tensor([width - 1, height - 1], dtype=self.dt) - 1
        return torch.cat([xy, side.reshape(rows, S).to(self.dt)], 1) if S else xy

    def attention(self, q, k, v, B, nq, nk, H, dh):
        d = H * dh
        qh = q.view(B, nq, H, dh).permute(0, 2, 1, 3)
        kh = k.view(B, nk, H, dh).permute(0, 2, 1, 3)
        vh = v.view(B, nk, H, dh).permute(0, 2, 1, 3)
        p = (qh @ kh.transpose(2, 3) * dh ** -0.5).softmax(-1)
        return (p @ vh).permute(0, 2, 1, 3).reshape(B * nq, d).contiguous()

    def softmax_rows(self, P, ld, rows, cols):
        v = self._view(P, 0, rows, cols, ld)
        v.copy_(v.softmax(-1))

    def softmax_bwd_rows(self, P, dP, ld, rows, cols, scale):
        p, g = self._view(P, 0, rows, cols, ld), self._view(dP, 0, rows, cols, ld)
        g.copy_(scale * p * (g - (p * g).sum(-1, keepdim=True)))

    def mix_fwd(self, g, l, mix):
        al = torch.sigmoid(mix.to(self.dt))
        return al * g + (1 - al) * l.to(self.dt)

    def mix_bwd(self, dm, mix):
        al = torch.sigmoid(mix.to(self.dt))
        return al * dm, (1 - al) * dm

    def mix_param_grad(self, csum, mix):
        al = torch.sigmoid(mix.to(self.dt))
        return csum * al * (1 - al)

    def bn_fwd(self, a, gamma, beta, eps, momentum, running_mean, running_var):
        r = a.clamp_min(0)
        mean, var = r.mean(0), r.var(0, unbiased=False)
        invstd = (var + eps).rsqrt()
        if running_mean is not None:
            n = a.shape[0]
            running_mean.mul_(1 - momentum).add_((momentum * mean).to(running_mean.dtype))
            running_var.mul_(1 - momentum).add_((momentum * var * n / max(n - 1, 1)).to(running_var.dtype))
        return (r - mean) * invstd * gamma.detach().to(self.dt) + beta.detach().to(self.dt), mean, invstd

    def bn_bwd(self, dy, a, gamma, mean, invstd):
        r = a.clamp_min(0)
        xhat = (r - mean) * invstd
        dbeta, dgamma = dy.sum(0), (dy * xhat).sum(0)
        n = a.shape[0]
        dr = gamma.detach().to(self.dt) * invstd * (dy - dbeta / n - xhat * dgamma / n)
        return dr * (a > 0), dgamma, dbeta

    def sinkhorn_fwd(self, Sp, dust, B, n, m, iters, reg):
        from oracle.sinkhorn_grad_oracle import forward_with_history
        return forward_with_history(Sp[:, :, :m].to(self.dt), dust.to(self.dt).reshape(()), iters, reg)[0], None

    def sinkhorn_bwd(self, Sp, dust, hist, G, B, n, m, iters, reg):
        from oracle.sinkhorn_grad_oracle import backward
        dS, dd = backward(Sp[:, :, :m].to(self.dt), dust.to(self.dt).reshape(()), iters, reg, G.to(self.dt))
        dZ = self.empty(B, n + 1, m + 1)                                 # only the inner block of d loss / d S_aug is consumed
        dZ[:, :n, :m] = dS
        return dZ, dd.reshape(1)

    def colsum(self, X, Y=None, Z=None):
        t = X if Y is None else X * (Y.to(self.dt) - (Z.to(self.dt) if Z is not None else 0))
        return t.sum(0)

    def axpby(self, x, y, a=1.0, b=1.0, out=None):
        r = a * x + (b * y if y is not None else 0)
        if out is not None:
            out.copy_(r)
            return out
        return r

    def sum_batches(self, part, S, rows, cols, out, out_off, ld, accumulate=True):
        t = part.reshape(S, rows, cols).sum(0)
        dst = self._view(out, out_off, rows, cols, ld)
        dst.copy_(dst + t if accumulate else t)

    # composite helpers: the product's own code, running on this double
    from openglue_b200.training import _Ops as _P
    linear, transpose, grad_weight, grad_input = _P.linear, _P.transpose, _P.grad_weight, _P.grad_input
    _transpose_chunks, SPLIT_K = _P._transpose_chunks, 32              # (small chunks: the fixtures have ~100-400 rows per image)


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_training_schedule_on_cpu_double(name):
    from openglue_b200 import SuperGlue
    from openglue_b200.training import TrainStep
    fx = _load(name)
    sd = synthetic_state_dict(fx['config'], seed=fx['weights_seed'])
    sd.update(fx['bn_buffers'])
    model = SuperGlue(dict(fx['config']))
    model.load_state_dict(sd, strict=True)
    model = model.double().train()
    data = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in fx['data'].items()}
    step = TrainStep(model, data, ops=_CpuOps())
    scores, c0, _ = step.forward()
    assert (scores - fx['scores_f64']).abs().max() < 5e-6               # (the reference keeps log_a / log_b in fp32 even in its fp64 run)
    assert (c0 - fx['context_descriptors0_f64']).abs().max() < 1e-9
    # d loss / d scores from the reference criterion's definition (the loss oracle)
    from oracle.loss_oracle import criterion_grad
    dscores = criterion_grad({'gt_matches0': fx['gt_matches0'], 'gt_matches1': fx['gt_matches1']}, tuple(scores.shape))
    grads = step.backward(dscores)
    for k, ref in fx['grads'].items():
        g = grads[k].reshape(ref.shape)
        assert torch.isfinite(g).all(), k
        assert (g - ref.double()).abs().max() <= 2e-6 * max(1.0, float(ref.abs().max())), k     # fixture gradients are stored in fp32
    for i in (0, 1):
        ref = fx[f'dlocal_descriptors{i}_f64']
        assert (grads[f'local_descriptors{i}'] - ref).abs().max() < 2e-6 * max(1.0, float(ref.abs().max()))
    for k, v in model.named_buffers():
        ref = fx['buffers_after'][k]
        assert (v.double() - ref.double()).abs().max() <= 1e-6 * max(1.0, float(ref.double().abs().max())), k


@pytest.mark.gpu
def test_training_step_pipeline_like_the_reference_module():
    """The reference's training_step with cached features (models/matching_module.py:71-105), every stage on the GPU library:
    labels (generate_gt_matches) -> SuperGlue in train() mode -> criterion -> nll_weight * loss + metric_weight * metric_loss ->
    backward -> optimiser.  The planted similarity of synthetic_pairs is the batch's ground-truth transformation."""
    from openglue_b200 import SuperGlue, criterion, generate_gt_matches
    from openglue_b200.synthetic import default_config, synthetic_pairs
    dev = torch.device('cuda:0')
    cfg = default_config(descriptor_dim=128, num_stages=2, num_iters=20)
    cfg['precision'] = 'tf32x3'
    model = SuperGlue(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=5), strict=True)
    model = model.to(dev).train()
    batch = 4
    pairs = synthetic_pairs(batch, 256, 300, 128, 1, family='planted', seed=21)
    pairs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in pairs.items()}
    H = torch.tensor([[0.9, 0.0, 20.0], [0.0, 0.9, 20.0], [0.0, 0.0, 1.0]], device=dev).repeat(batch, 1, 1)
    raw = {'transformation': {'type': ['perspective'] * batch, 'H': H}, 'image0_size': pairs['image0_size'], 'image1_size': pairs['image1_size']}
    f0 = {'keypoints': pairs['keypoints0'], 'side_info': pairs['side_info0'], 'local_descriptors': pairs['local_descriptors0']}
    f1 = {'keypoints': pairs['keypoints1'], 'side_info': pairs['side_info1'], 'local_descriptors': pairs['local_descriptors1']}
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    losses = []
    for _ in range(4):
        data, y_true = generate_gt_matches(raw, f0, f1, 3.0, 5.0)
        assert int((y_true['gt_matches0'] >= 0).sum()) > 0
        y_pred = model(data)
        loss = criterion(y_true, y_pred, margin=None)
        total = 1.0 * loss['loss'] + 0.0 * loss['metric_loss']
        opt.zero_grad()
        total.backward()
        for k, p in model.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
        opt.step()
        losses.append(float(total.detach()))
    assert losses[-1] < losses[0], losses


@pytest.mark.gpu
def test_graphed_training_step_is_bit_identical_to_the_eager_step():
    """GraphedTrainStep (the whole step replayed as one CUDA graph) against the eager autograd route: same loss, same gradients,
    same BatchNorm buffers, bit for bit, over several optimiser steps."""
    import copy
    from openglue_b200 import SuperGlue, criterion
    from openglue_b200.training import GraphedTrainStep
    dev = torch.device('cuda:0')
    fx = _load('train_ragged')
    sd = synthetic_state_dict(fx['config'], seed=fx['weights_seed'])
    sd.update(fx['bn_buffers'])
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in fx['data'].items()}
    y_true = {'gt_matches0': fx['gt_matches0'].to(dev), 'gt_matches1': fx['gt_matches1'].to(dev)}
    models = []
    for _ in range(2):
        m = SuperGlue(dict(fx['config'], precision='tf32x3'))
        m.load_state_dict(copy.deepcopy(sd), strict=True)
        models.append(m.to(dev).train())
    eager, graphed = models
    opt_e = torch.optim.SGD(eager.parameters(), lr=1e-3)
    opt_g = torch.optim.SGD(graphed.parameters(), lr=1e-3)
    step = GraphedTrainStep(graphed, data, y_true)
    for it in range(3):
        opt_e.zero_grad()
        loss_e = criterion(y_true, eager(data), margin=None)['loss']
        loss_e.backward()
        loss_g = step(data, y_true)['loss']
        assert torch.equal(loss_e.detach(), loss_g), it
        for (k, pe), (_, pg) in zip(eager.named_parameters(), graphed.named_parameters()):
            assert torch.equal(pe.grad, pg.grad), (it, k)
        opt_e.step()
        opt_g.step()
    for (k, be), (_, bg) in zip(eager.named_buffers(), graphed.named_buffers()):
        assert torch.equal(be, bg), k
