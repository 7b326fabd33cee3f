#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_attn2 -f python scripts/prof_ops.py attn 2 > gpurun_out/prof_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc2 -s 3 -c 3 -o gpurun_out/prof_linear2 -f python scripts/prof_ops.py linear 2 > gpurun_out/prof_linear.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn -s 1 -c 1 -o gpurun_out/prof_sinkhorn2 -f python scripts/prof_ops.py sinkhorn 2 > gpurun_out/prof_sinkhorn.log 2>&1
ls -la gpurun_out/*2.ncu-rep
