#!/bin/bash
# ground-truth match generation (row f2): GPU parity tests + timing against the oracle on the host cores
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gt_matches.py -x -q -m gpu > gpurun_out/gt_tests.log 2>&1; echo "rc=$?" >> gpurun_out/gt_tests.log
tail -4 gpurun_out/gt_tests.log
timeout 300 python - > gpurun_out/gt_timing.txt 2>&1 <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
from openglue_b200.synthetic import synthetic_gt_scene
from openglue_b200.gt_matches import gt_matches
from oracle import gt_matches_oracle as G
for kind in ('perspective', '3d_reprojection'):
    sc = synthetic_gt_scene(16, 2048, 2048, kind, seed=3)
    dev = torch.device('cuda')
    k0, k1 = sc['keypoints0'].to(dev), sc['keypoints1'].to(dev)
    tf = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc['transformation'].items()}
    for _ in range(3): gt_matches(k0, k1, tf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): gt_matches(k0, k1, tf)
    e1.record(); torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 20
    best = 1e9
    for th in (8, 16, 32):
        torch.set_num_threads(th)
        G.gt_matches(sc['keypoints0'], sc['keypoints1'], sc['transformation'])
        t0 = time.perf_counter(); G.gt_matches(sc['keypoints0'], sc['keypoints1'], sc['transformation']); best = min(best, (time.perf_counter() - t0) * 1e3)
    print(f'{kind}: 16 pairs x 2048 x 2048  GPU {gpu_ms:.3f} ms per batch (7 launches, python call included)  oracle (torch CPU, best of 8/16/32 threads) {best:.1f} ms')
PY
cat gpurun_out/gt_timing.txt | tail -4
