#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_attn7 -f python scripts/prof_ops.py attn 2 > gpurun_out/prof_attn.log 2>&1
ls -la gpurun_out/prof_attn7.ncu-rep
