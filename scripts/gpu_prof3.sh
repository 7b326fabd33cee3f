#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_attn3 -f python scripts/prof_ops.py attn 2 > gpurun_out/prof_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc2 -s 3 -c 3 -o gpurun_out/prof_linear3 -f python scripts/prof_ops.py linear 2 > gpurun_out/prof_linear.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_tc.log 2>&1
ls -la gpurun_out/*3.ncu-rep
