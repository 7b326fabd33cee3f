#!/bin/bash
# visit J: Sinkhorn L2-resident subset sweep
mkdir -p gpurun_out
for mb in 0 32 48 64 80 96 112; do
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | tail -1
done | tee gpurun_out/j_sink_l2_C3.txt
for mb in 0 48 64 80 96 112 140; do
  OG_SINK_L2_MB=$mb timeout 120 python scripts/sink_l2_exp.py 32 1024 1024 100 2>&1 | tail -1
done | tee gpurun_out/j_sink_l2_C2.txt
for mb in 0 64 96; do
  OG_SINK_L2_MB=$mb timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/j_bench_l2_$mb.json 2> gpurun_out/j_bench_l2_$mb.err
  python - <<P
import json
d=[json.loads(l) for l in open('gpurun_out/j_bench_l2_$mb.json') if l.startswith('{')][-1]
print('bench L2_MB=$mb', d['value'], d['ms_per_step'], d['roofline_sinkhorn']['ms_per_launch'], d['roofline']['achieved'])
P
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sinkhorn or oracle" 2>&1 | tail -3
