#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 1500 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_tc.log
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/tc_probe.log; tail -12 gpurun_out/pytest_tc.log
