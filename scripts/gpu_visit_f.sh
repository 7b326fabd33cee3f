#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
tail -3 gpurun_out/f_tests.log
timeout 300 python bench.py --steps 8 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/f_bench.json'))
print(round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), ' attn', round(d['roofline']['ms_per_launch'], 4), 'ms  sink', round(d['roofline_sinkhorn']['ms_per_launch'], 3), d['clocks'])
PY
OG_ATTN_PAIR=0 timeout 300 python bench.py --steps 4 --no-cpu-baseline > gpurun_out/f_bench_single.json 2> gpurun_out/f_bench_single.err
python - <<PY
import json
d = json.load(open('gpurun_out/f_bench_single.json'))
print('attn single-CTA form:', round(d['value'], 1), 'pairs/s', ' attn', round(d['roofline']['ms_per_launch'], 4))
PY
timeout 300 python bench.py --workload C5 --steps 8 --no-cpu-baseline > gpurun_out/f_bench_C5.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/f_bench_C5.json')); print('C5', round(d['value'],1), round(d['ms_per_step'],3))"
OG_ATTN_PAIR=1 timeout 120 python scripts/trace_attn.py > gpurun_out/trace_attn_pair3.log 2>&1
