#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 600 -k "attention_tc" 2>&1 | tail -30 > gpurun_out/pytest_tc_attn.log
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 600 -k "not attention_tc" 2>&1 | tail -30 > gpurun_out/pytest_tc.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32x3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
tail -15 gpurun_out/pytest_tc_attn.log; tail -8 gpurun_out/pytest_tc.log; head -c 2500 gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
