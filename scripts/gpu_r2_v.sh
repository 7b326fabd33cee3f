#!/bin/bash
# Round-2 visit V: clean full GPU test run of the committed tree + SuperPoint timing.
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 | tee gpurun_out/v_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/v_smoke.log
timeout 300 python - <<'P' 2>&1 | tee gpurun_out/v_superpoint_time.txt
import sys, time, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
from gen_golden_superpoint import synthetic_superpoint_state_dict, synthetic_images
from openglue_b200 import SuperPointNet, _cabi
dev = torch.device('cuda:0')
for prec in ('tf32x3', 'fp32'):
    m = SuperPointNet(max_keypoints=2048, precision=prec); m.load_state_dict(synthetic_superpoint_state_dict(1)); m = m.to(dev).eval()
    img = synthetic_images(2, 480, 640, 5).to(dev)
    for _ in range(2): out = m(img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = m(img)
    e1.record(); torch.cuda.synchronize()
    print('SuperPointNet %s: 2 images 480x640, max_keypoints 2048: %.2f ms per image, %d keypoints' % (prec, e0.elapsed_time(e1) / 10, out[0].shape[1]))
P
