#!/bin/bash
# Round-2 visit G: fp16 GEMM with two converter teams (parity, timing, trace), ncu of the default fp16 attention, other workloads.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -s --timeout 500 2>&1 | grep -v "^$" | tail -40 > gpurun_out/g_$name.log; echo "== $name"; grep -E "^\[|passed|failed|rror" gpurun_out/g_$name.log | tail -8; }
run gemm tests/test_gpu_f16.py -k "linear_f16 or layout_probe"
run path tests/test_gpu_parity.py -k "(reference_big and fp16x3-1) or (matches_oracle and fp16x3)"
timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 8 > gpurun_out/g_bench_f16.json 2> gpurun_out/g_bench_f16.err
for wl in C1 C2 C5; do
  timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 8 --workload $wl > gpurun_out/g_bench_f16_$wl.json 2> gpurun_out/g_bench_f16_$wl.err
done
python - <<'PY'
import json
for n in ('g_bench_f16','g_bench_f16_C1','g_bench_f16_C2','g_bench_f16_C5'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],3), 'ms; attn', round(d['roofline']['ms_per_launch'],4), 'ms', round(d['roofline']['achieved'],1), 'TF/s; sinkhorn', round(d['roofline_sinkhorn']['frac'],3), 'verified', (d.get('verified') or {}).get('matches0_identical'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-400:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/g_launches_f16.csv python bench.py --precision fp16x3 --steps 1 --warmup 1 --no-cpu-baseline --no-verify --cuda-graph 0 > gpurun_out/g_ncu_bench.log 2>&1
python scripts/agg_launches.py gpurun_out/g_launches_f16.csv > gpurun_out/g_launches_f16_agg.txt 2>&1; head -12 gpurun_out/g_launches_f16_agg.txt
timeout 200 python scripts/trace_f16.py gemm fc2 > gpurun_out/g_trace_gemm_f16_fc2.log 2>&1; sed -n 1,26p gpurun_out/g_trace_gemm_f16_fc2.log; tail -9 gpurun_out/g_trace_gemm_f16_fc2.log
