#!/bin/bash
set +e
mkdir -p gpurun_out
for cl in 1 2 4; do
  echo "== cluster $cl" >> gpurun_out/tc_probe_cl.log
  OG_TC_CLUSTER=$cl timeout 600 python scripts/tc_probe.py v2_tails_concat v2_resid_split_t v2_batchedB v2_qkv_shape_time v2_fc1_shape_time v2_fc2_shape_time v2_score_shape_time >> gpurun_out/tc_probe_cl.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_tc.log
OG_TC_CLUSTER=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl2.json 2> gpurun_out/bench_cl2.err
OG_TC_CLUSTER=4 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl4.json 2> gpurun_out/bench_cl4.err
cat gpurun_out/tc_probe_cl.log; tail -4 gpurun_out/pytest_tc.log; head -c 250 gpurun_out/bench_cl2.json; echo; head -c 250 gpurun_out/bench_cl4.json; tail -3 gpurun_out/bench_cl2.err gpurun_out/bench_cl4.err
