#!/bin/bash
# ncu evidence with the current defaults (cta_group::2 kernels): launch list of one bench step + full captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_attn5 -f python scripts/prof_ops.py attn 2 > gpurun_out/prof_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc2 -s 3 -c 3 -o gpurun_out/prof_linear4 -f python scripts/prof_ops.py linear 2 > gpurun_out/prof_linear.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn -s 1 -c 1 -o gpurun_out/prof_sinkhorn3 -f python scripts/prof_ops.py sinkhorn 2 > gpurun_out/prof_sinkhorn.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
