#!/bin/bash
# visit K: Sinkhorn L2-resident subset with a persisting-L2 carve-out
mkdir -p gpurun_out
for mb in 0 32 48 64 80; do
  OG_PERSIST_MB=200 OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | tail -2
done | tee gpurun_out/k_sink_l2_C3.txt
for mb in 48 80; do
  OG_PERSIST_MB=64 OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | tail -2
done | tee -a gpurun_out/k_sink_l2_C3.txt
for mb in 0 48 64 80 140; do
  OG_PERSIST_MB=200 OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 32 1024 1024 100 2>&1 | tail -1
done | tee gpurun_out/k_sink_l2_C2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "oracle and flat" 2>&1 | tail -40 > gpurun_out/k_test.txt
tail -5 gpurun_out/k_test.txt
