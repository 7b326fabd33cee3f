#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/d_tests.log
tail -3 gpurun_out/d_tests.log
timeout 300 python bench.py --steps 8 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/d_bench.json'))
print(round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), ' attn', round(d['roofline']['ms_per_launch'], 4), 'ms  sink', round(d['roofline_sinkhorn']['ms_per_launch'], 3), round(d['roofline_sinkhorn']['frac'], 3), d['clocks'])
PY
