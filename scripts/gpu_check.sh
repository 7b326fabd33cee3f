#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Outputs -> gpurun_out/
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > gpurun_out/cpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ -n "$WITH_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-400} --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json | head -c 1500
