#!/bin/bash
# split-QK attention: parity, bench, trace; A/B of experimental builds (OG_LIB)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c_tests.log
tail -3 gpurun_out/c_tests.log
run_bench() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --no-cpu-baseline > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/c_bench_$name.json'))
    print('$name', round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms  attn', round(d['roofline']['ms_per_launch'], 4), 'ms  sink', round(d['roofline_sinkhorn']['ms_per_launch'], 3))
except Exception as e:
    print('$name failed', e)
PY
}
run_bench base OG_X=0
for v in openglue_b200/variants/lib_*.so; do
  n=$(basename $v .so)
  OG_LIB=$PWD/$v timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/c_tests_$n.log 2>&1; echo "rc=$?" >> gpurun_out/c_tests_$n.log
  tail -2 gpurun_out/c_tests_$n.log
  run_bench $n OG_LIB=$PWD/$v
done
OG_ATTN_PAIR=1 timeout 120 python scripts/trace_attn.py > gpurun_out/trace_attn_pair2.log 2>&1
