#!/bin/bash
# visit N: Sinkhorn with packed f32x2 sweep + unmasked fast path, L2-resident subset sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sinkhorn_grad.py -x -q -m gpu -k "sinkhorn or oracle" 2>&1 | tail -3
for mb in 0 32 48 64 80 96; do
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | tail -1
done | tee gpurun_out/n_sink_l2_C3.txt
for mb in 0 64 96 112; do
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 32 1024 1024 100 2>&1 | tail -1
done | tee gpurun_out/n_sink_l2_C2.txt
for mb in 0 64; do
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 1 512 512 20 2>&1 | tail -1
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 1 4096 1024 50 2>&1 | tail -1
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 4 3000 4000 100 2>&1 | tail -1
done | tee gpurun_out/n_sink_l2_misc.txt
