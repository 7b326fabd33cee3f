#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32x3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
timeout 600 python scripts/tc_probe.py v2_qkv_shape_time v2_fc1_shape_time v2_fc2_shape_time v2_score_shape_time > gpurun_out/tc_probe.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision tf32x3 > gpurun_out/ncu_bench_tc.log 2>&1
tail -8 gpurun_out/pytest_all.log; cat gpurun_out/tc_probe.log; head -c 400 gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
