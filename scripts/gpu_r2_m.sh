#!/bin/bash
# visit M: Sinkhorn with an INTERLEAVED L2-resident subset
mkdir -p gpurun_out
for mb in 0 32 48 64 80 96; do
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | tail -1
done | tee gpurun_out/m_sink_l2_C3.txt
for mb in 0 64 96 112; do
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 32 1024 1024 100 2>&1 | tail -1
done | tee gpurun_out/m_sink_l2_C2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "oracle" 2>&1 | tail -3
