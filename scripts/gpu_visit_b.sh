#!/bin/bash
# attention event traces (both kernel forms) + the other BASELINE.json configurations with the current defaults
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
OG_ATTN_PAIR=1 timeout 120 python scripts/trace_attn.py > gpurun_out/trace_attn_pair.log 2>&1
OG_ATTN_PAIR=0 timeout 120 python scripts/trace_attn.py > gpurun_out/trace_attn_single.log 2>&1
for w in C1 C2 C5; do
  timeout 300 python bench.py --workload $w --steps 8 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value'],1))"
done
