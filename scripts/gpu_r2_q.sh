#!/bin/bash
# visit Q2: Sinkhorn with an L2 prefetch ahead of the smem ring
mkdir -p gpurun_out
for pf in 0 1 2 4 8; do
  OG_SINK_PF=$pf timeout 120 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | tail -1 | sed "s/^/PF=$pf /"
done | tee gpurun_out/q_sink_pf_C3.txt
for pf in 2 4; do
  OG_SINK_L2_MB=64 OG_SINK_PF=$pf timeout 120 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | tail -1 | sed "s/^/PF=$pf /"
done | tee -a gpurun_out/q_sink_pf_C3.txt
for pf in 0 2 4 8; do
  OG_SINK_PF=$pf timeout 120 python scripts/sink_l2_exp.py 32 1024 1024 100 2>&1 | tail -1 | sed "s/^/PF=$pf /"
done | tee gpurun_out/q_sink_pf_C2.txt
for pf in 0 4; do
  OG_SINK_PF=$pf timeout 120 python scripts/sink_l2_exp.py 1 512 512 20 2>&1 | tail -1 | sed "s/^/PF=$pf /"
  OG_SINK_PF=$pf timeout 120 python scripts/sink_l2_exp.py 1 4096 1024 50 2>&1 | tail -1 | sed "s/^/PF=$pf /"
done | tee gpurun_out/q_sink_pf_misc.txt
