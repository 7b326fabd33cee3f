#!/bin/bash
# Round-2 visit I: fp16 GEMM with separate A / B rings (parity, timing, trace); collation tests.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -s --timeout 500 2>&1 | grep -v "^$" | tail -40 > gpurun_out/i_$name.log; echo "== $name"; grep -E "^\[|passed|failed|rror" gpurun_out/i_$name.log | tail -8; }
run gemm tests/test_gpu_f16.py -k "linear_f16 or layout_probe"
run collate tests/test_collate.py
run path tests/test_gpu_parity.py -k "(reference_big and fp16x3) or (matches_oracle and fp16x3) or no_descriptors"
timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload C2 > gpurun_out/i_bench_C2.json 2> gpurun_out/i_bench_C2.err
python - <<'PY'
import json
for n in ('i_bench','i_bench_C2'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'verified', (d.get('verified') or {}).get('matches0_identical'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-400:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify --cuda-graph 0 > gpurun_out/i_ncu_bench.log 2>&1
python scripts/agg_launches.py gpurun_out/i_launches.csv > gpurun_out/i_launches_agg.txt 2>&1; head -9 gpurun_out/i_launches_agg.txt
timeout 200 python scripts/trace_f16.py gemm fc2 > gpurun_out/i_trace_gemm_f16_fc2.log 2>&1; sed -n 1,22p gpurun_out/i_trace_gemm_f16_fc2.log; tail -9 gpurun_out/i_trace_gemm_f16_fc2.log
timeout 200 python scripts/trace_f16.py gemm q > gpurun_out/i_trace_gemm_f16_q.log 2>&1; tail -9 gpurun_out/i_trace_gemm_f16_q.log
