#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 300 python -m pytest tests/test_training.py -x -q -m gpu -k "pipeline or ragged-tf32x3" 2>&1 | grep -E "ERROR SUMMARY|Invalid|at 0x|passed|failed" | head -20
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu -k "not big" 2>&1 | tail -3
