#!/bin/bash
# A/B timing of experimental builds (openglue_b200/variants/lib_*.so, selected with OG_LIB) against the default library
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run_bench() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_$name.json'))
    print('$name', round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms  attn', round(d['roofline']['ms_per_launch'], 4), 'ms', d['clocks']['sm_mhz'])
except Exception as e:
    print('$name failed', e)
PY
}
run_bench base_1 OG_X=0
for v in openglue_b200/variants/lib_*.so; do
  n=$(basename $v .so)
  OG_LIB=$PWD/$v timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/ab_tests_$n.log 2>&1; echo "rc=$?" >> gpurun_out/ab_tests_$n.log
  tail -2 gpurun_out/ab_tests_$n.log
  run_bench ${n}_1 OG_LIB=$PWD/$v
done
run_bench base_2 OG_X=0
for v in openglue_b200/variants/lib_*.so; do n=$(basename $v .so); run_bench ${n}_2 OG_LIB=$PWD/$v; done
