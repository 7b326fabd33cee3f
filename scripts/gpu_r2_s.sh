#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_f16.py tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/dbg_fuse.py 2>&1 | grep -v fp32 | tail -12
