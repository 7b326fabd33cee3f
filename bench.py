#!/usr/bin/env python
"""bench.py - image-pairs/sec of the OpenGlue matching core on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the whole hot path (keypoint encoder -> 18 attention layers -> score
matrix -> 100 Sinkhorn iterations -> mutual matches) over one batch of synthetic image pairs per
GPU (weak scaling: per-GPU batch fixed, pairs are independent, no data-path collective; one tiny
NCCL all-reduce of the per-rank match statistics per step mirrors the reference's
`self.log(..., sync_dist=True)`).

Prints ONE JSON line (rank 0).  Keys follow the driver's contract; `roofline` is measured live
on the dominant kernel through the operator-level C-ABI call, `cpu_baseline` / `--impl reference`
time the oracle port (same ATen ops as the reference's PyTorch-CPU path) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from openglue_b200.synthetic import BASELINE_CONFIGS, default_config, synthetic_pairs, synthetic_state_dict  # noqa: E402

METRIC = 'image-pairs/sec at N=M=2048, d=256, 9 GNN layers, 100 Sinkhorn iters'
MATCH_THRESHOLD = 0.2


def flops_per_pair(n, m, d, stages, s):
    """SURVEY.md section 8(d) / BASELINE.md section 4 (multiply-add = 2 FLOP, exp not counted)."""
    f_attn = 4 * d * stages * (n + m) ** 2
    f_lin = 40 * stages * (n + m) * d * d
    f_final = 2 * (n + m) * d * d + 2 * n * m * d
    f_pe = 2 * (n + m) * (32 * (2 + s) + 32 * 64 + 64 * 128 + 128 * d)
    return dict(total=f_attn + f_lin + f_final + f_pe, attn=f_attn, lin=f_lin)


def sinkhorn_bytes_per_pair(n, m, iters):
    return (iters + 1) * 4 * (n + 1) * (m + 1)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], bf16_tflops=p['bf16_tflops'],
                    bf16_tflops_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']), source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        busy = sorted(sm)[len(sm) // 2:]                      # upper half = samples under load
        return {'sm_mhz': statistics.median(busy), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def time_oracle(cfg, n, m, reps, family='planted'):
    """The CPU path of the reference (oracle port: same ATen ops), B = 1 pair per run.
    torch's intra-op thread count is swept first (on a 2-socket host more threads is not faster);
    the timed runs use the best count.  Returns (times, threads)."""
    from oracle import superglue_oracle as O                  # checker / CPU baseline only
    sd = synthetic_state_dict(cfg, seed=0)
    data = synthetic_pairs(1, n, m, cfg['descriptor_dim'], cfg['positional_encoding']['side_info_size'],
                           family=family, seed=1234)
    default_threads = torch.get_num_threads()             # torchrun pins OMP_NUM_THREADS=1: sweep by core count instead
    ncpu = os.cpu_count() or 8
    cands = sorted({t for t in (8, 16, 32, 64, default_threads) if t <= ncpu})
    best_t, best = default_threads, float('inf')
    for t in cands:
        torch.set_num_threads(t)
        O.run(sd, cfg, data, MATCH_THRESHOLD)                 # warm-up at this thread count
        t0 = time.perf_counter()
        O.run(sd, cfg, data, MATCH_THRESHOLD)
        dt = time.perf_counter() - t0
        if dt < best:
            best_t, best = t, dt
        if dt > 4 * best:                                     # clearly past the knee; stop sweeping
            break
    torch.set_num_threads(best_t)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.run(sd, cfg, data, MATCH_THRESHOLD)
        times.append(time.perf_counter() - t0)
    torch.set_num_threads(default_threads)
    return times, best_t


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on the host cores
    (oracle port; /root/reference is Python and does not travel to the GPU box)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = default_config(**wl['cfg'])
    # each step = a bounded sample of the workload: ONE pair of the workload's shape
    times, threads = time_oracle(cfg, wl['n'], wl['m'], max(1, args.steps))
    sec = sum(times) / len(times)
    value = 1.0 / sec
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': 1, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': bench_config(args, wl, per_gpu_batch=1),
        'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                         'sample': f'1 pair per step of the {args.workload} shape (N={wl["n"]}, M={wl["m"]}), '
                                   f'{len(times)} timed runs after warm-up, torch CPU fp32, {threads} threads (best of a sweep), '
                                   f'os.cpu_count()={os.cpu_count()}'},
        'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(args, wl, per_gpu_batch):
    c = wl['cfg']
    return {'workload': f'{args.workload}: {per_gpu_batch} pairs/GPU, N={wl["n"]}, M={wl["m"]}, d={c["descriptor_dim"]}, '
                        f'{c["num_stages"]} stages (self+cross), {c["num_iters"]} Sinkhorn iters, planted synthetic pairs',
            'pairs_per_gpu': per_gpu_batch, 'global_pairs': per_gpu_batch * args.gpus,
            'N': wl['n'], 'M': wl['m'], 'd': c['descriptor_dim'], 'stages': c['num_stages'],
            'sinkhorn_iters': c['num_iters'], 'parallelism': f'pairs sharded over {args.gpus} GPU(s), no data-path collective',
            'l2': 'working set per step (scores 16.8 MB/pair + activations) exceeds the 126 MB L2; no explicit flush',
            'cuda_graph': bool(getattr(args, 'cuda_graph', 0))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='C3', choices=sorted(BASELINE_CONFIGS))
    ap.add_argument('--pairs-per-gpu', type=int, default=None)
    ap.add_argument('--precision', default=os.environ.get('OG_PRECISION', 'tf32x3'), choices=['fp32', 'tf32x3'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cuda-graph', type=int, default=1, help='replay the launch schedule from a CUDA graph (default on)')
    args = ap.parse_args()
    wl = dict(BASELINE_CONFIGS[args.workload])
    if args.impl == 'reference':
        return run_reference(args, wl)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from openglue_b200 import _cabi
    from openglue_b200.build import build
    if rank == 0:
        build()
    if dist is not None:
        dist.barrier()
    from openglue_b200.superglue import MatchingCore, SuperGlue

    batch = args.pairs_per_gpu or wl['batch']
    n, m = wl['n'], wl['m']
    cfg = default_config(**wl['cfg'])
    cfg['precision'] = args.precision
    d, s_dim, stages, iters = cfg['descriptor_dim'], cfg['positional_encoding']['side_info_size'], \
        cfg['attention_gnn']['num_stages'], cfg['otp']['num_iters']
    model = SuperGlue(cfg).eval()
    model.load_state_dict(synthetic_state_dict(cfg, seed=0))
    model = model.to(dev)
    core = MatchingCore(model, MATCH_THRESHOLD, device=dev, use_cuda_graph=bool(args.cuda_graph))
    host = synthetic_pairs(batch, n, m, d, s_dim, family='planted', seed=1234 + rank)
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    stats = torch.zeros(2, device=dev)

    from openglue_b200.sharding import all_reduce_statistics, match_statistics

    def step(inputs):
        res = core(inputs)
        if dist is not None:                  # the reference's sync_dist logging: one tiny NCCL all-reduce per step
            st = match_statistics(res['matches0'], res['matching_scores0']).to(dev)
            dist.all_reduce(st)
            stats.copy_(st[:2])
        return res

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput (value) ----
    for _ in range(max(3, args.warmup)):
        step(data)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: step(data), args.steps)
    launches = model.last_launches * args.steps
    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region) ----
    for _ in range(2):
        step(host)
    ms_e2e_serial = timed(lambda: step(host), args.steps)          # blocking forward(): copy, compute, copy, in series

    # serving form of the same API: submit()/wait() with two batches in flight, so the upload of step i+1 overlaps
    # the kernels of step i; every step still uploads its inputs and reads its matches back inside the timed region
    def consume(res):
        if dist is not None:
            st = match_statistics(res['matches0'], res['matching_scores0']).to(dev)
            dist.all_reduce(st)
            stats.copy_(st[:2])

    def pipelined(steps):
        pend = None
        for _ in range(steps):
            nxt = core.submit(host)
            if pend is not None:
                consume(pend.wait())
            pend = nxt
        consume(pend.wait())
    pipelined(2)
    ms_e2e = timed(lambda: pipelined(args.steps), 1)
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(v.numel() * v.element_size() for k, v in host.items() if torch.is_tensor(v) and k in core._TENSOR_KEYS)
    d2h = batch * (n * 8 + n * 4 + m * 8 + m * 4)

    ms_per_step = ms_total / args.steps
    value = batch * world / (ms_per_step * 1e-3)
    e2e_value = batch * world / (ms_e2e / args.steps * 1e-3)

    # ---- roofline of the dominant kernel: fused attention (self layer launch), measured live ----
    peaks = measured_peaks()
    lib = _cabi.lib()
    import ctypes as C
    H = cfg['attention_gnn']['num_heads']
    nb = 2 * batch if n == m else batch
    qkv = torch.randn(nb * n, 3 * d, device=dev)
    o = torch.empty(nb * n, d, device=dev)
    prec = {'fp32': _cabi.OG_PREC_FP32, 'tf32x3': _cabi.OG_PREC_TF32X3}[args.precision]
    p = lambda t, off=0: C.c_void_p(t.data_ptr() + off * 4)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    if args.precision == 'tf32x3':
        kk = torch.randn(nb * n, d, device=dev)
        ldv = (n + 3) // 4 * 4
        vt = torch.randn(nb * d, ldv, device=dev)
        khi, klo, vthi, vtlo = (torch.empty_like(t) for t in (kk, kk, vt, vt))
        _cabi.check(lib.og_split_tf32(p(kk), p(khi), p(klo), kk.numel(), st), 'og_split_tf32')
        _cabi.check(lib.og_split_tf32(p(vt), p(vthi), p(vtlo), vt.numel(), st), 'og_split_tf32')
        qq = torch.randn(nb * n, d, device=dev)

        def attn():
            _cabi.check(lib.og_attention_tc_fwd(p(qq), d, n * d, p(khi), p(klo), d, p(vthi), p(vtlo), ldv, p(o), d, n * d,
                                                nb, n, n, H, d // H, st), 'og_attention_tc_fwd')
    else:
        def attn():
            _cabi.check(lib.og_attention_fwd(p(qkv), 3 * d, n * 3 * d, p(qkv, d), 3 * d, n * 3 * d, p(qkv, 2 * d), 3 * d,
                                             n * 3 * d, p(o), d, n * d, nb, n, n, H, d // H, prec, st), 'og_attention_fwd')
    for _ in range(3):
        attn()
    reps = 10
    ms_attn = timed(attn, reps) / reps
    attn_flops = 4.0 * n * n * d * nb
    attn_tflops = attn_flops / (ms_attn * 1e-3) / 1e12
    # secondary: the Sinkhorn kernel against the HBM roofline
    lds = (m + 3) // 4 * 4
    sbuf = torch.randn(batch, n, lds, device=dev) * 4
    sc = torch.empty(batch, n + 1, m + 1, device=dev)
    wsb = lib.og_sinkhorn_workspace_bytes(batch, n, m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    dust = torch.ones(1, device=dev)

    def sink():
        _cabi.check(lib.og_sinkhorn_fwd(p(sbuf), lds, n * lds, p(dust), batch, n, m, iters, 1.0, p(sc), p(ws), wsb, st),
                    'og_sinkhorn_fwd')
    for _ in range(2):
        sink()
    ms_sink = timed(sink, 5) / 5
    sink_gbs = sinkhorn_bytes_per_pair(n, m, iters) * batch / (ms_sink * 1e-3) / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    if os.path.exists(tpath) and args.workload == 'C3' and batch == 16 and args.precision == 'tf32x3':
        traffic = json.load(open(tpath))     # ncu dram bytes per launch, captured at exactly these shapes
    fl = flops_per_pair(n, m, d, stages, s_dim)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'tf32x3 (fp32 accumulate)',
        'data': 'synthetic', 'config': bench_config(args, wl, batch),
        'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps, 'api': 'MatchingCore.submit()/wait(), host buffers, 2 batches in flight',
                'blocking_forward_value': batch * world / (ms_e2e_serial / args.steps * 1e-3),
                'blocking_forward_ms_per_step': ms_e2e_serial / args.steps},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': {'kernel': 'fused attention (self layer: %d sequences x %d heads, %d x %d, Dh=%d)' % (nb, H, n, n, d // H),
                     'bound': 'tensor', 'achieved': attn_tflops, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                     'frac': attn_tflops / peaks['bf16_tflops'],
                     'traffic': traffic.get('attention_tc_self_32seq_2048', {}).get('bytes'), 'peak_source': peaks['source'] + ' bf16 burst',
                     'ms_per_launch': ms_attn, 'flops_per_launch': attn_flops,
                     # the kernel runs 3 tf32 MMAs per algorithmic product (fp32-grade accuracy is part of the contract);
                     # tf32 dense rate = half the bf16 rate, so its own ceiling is bf16_peak / 6
                     'frac_of_3xtf32_ceiling': attn_tflops / (peaks['bf16_tflops'] / 6.0) if args.precision == 'tf32x3' else None},
        'roofline_sinkhorn': {'kernel': 'sinkhorn (%d pairs, %d iterations, one launch)' % (batch, iters), 'bound': 'hbm',
                              'achieved': sink_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                              'frac': sink_gbs / peaks['hbm_gbs'],
                              'traffic': traffic.get('sinkhorn_16pairs_2048_100it', {}).get('bytes'), 'peak_source': peaks['source'],
                              'ms_per_launch': ms_sink, 'bytes_per_launch': sinkhorn_bytes_per_pair(n, m, iters) * batch},
        'flops_per_pair': fl['total'],
        'end_to_end_tensor_frac': value / world * fl['total'] / (peaks['bf16_tflops_sustained'] * 1e12),
    }
    if world == 1 and not args.no_cpu_baseline:
        reps_cpu = 3
        times, threads = time_oracle(default_config(**wl['cfg']), n, m, reps_cpu)
        line['cpu_baseline'] = {'value': 1.0 / min(times), 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                                'sample': f'{reps_cpu} single pairs of the {args.workload} shape (N={n}, M={m}) after 1 '
                                          f'warm-up, best run, torch CPU fp32, {threads} threads (best of a sweep), '
                                          f'os.cpu_count()={os.cpu_count()}'}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
