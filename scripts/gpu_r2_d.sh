#!/bin/bash
# Round-2 visit D: two-team fp16 attention kernel (parity, timing against the one-team form, event trace), fp16 GEMM with K = 128
# accumulator chunks, Sinkhorn configuration experiments.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -s --timeout 500 2>&1 | grep -v "^$" | tail -40 > gpurun_out/d_$name.log; echo "== $name"; grep -E "^\[|passed|failed|error|Error" gpurun_out/d_$name.log | tail -14; }
run attn tests/test_gpu_f16.py -k "attention_f16"
run gemm tests/test_gpu_f16.py -k "linear_f16"
run path tests/test_gpu_parity.py -k "(reference_big and fp16x3) or (matches_oracle and fp16x3)"
for teams in 1 0; do
  OG_ATTN_TEAMS=$teams timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 6 > gpurun_out/d_bench_f16_teams$teams.json 2> gpurun_out/d_bench_f16_teams$teams.err
done
for cfg in 824 1612 444; do
  OG_SINK_CFG=$cfg timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 4 > gpurun_out/d_bench_f16_sink$cfg.json 2> gpurun_out/d_bench_f16_sink$cfg.err
done
python - <<'PY'
import json
for n in ('d_bench_f16_teams1','d_bench_f16_teams0','d_bench_f16_sink824','d_bench_f16_sink1612','d_bench_f16_sink444'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms; attn', round(d['roofline']['ms_per_launch'],4), 'ms', round(d['roofline']['achieved'],1), 'TF/s; sinkhorn', round(d['roofline_sinkhorn']['ms_per_launch'],3), 'ms; verified', d['verified']['matches0_identical'])
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-600:])
PY
timeout 200 python scripts/trace_f16.py attn > gpurun_out/d_trace_attn_f16t.log 2>&1; head -30 gpurun_out/d_trace_attn_f16t.log
timeout 200 python scripts/trace_f16.py gemm fc2 > gpurun_out/d_trace_gemm_f16_fc2.log 2>&1; sed -n 1,30p gpurun_out/d_trace_gemm_f16_fc2.log; tail -9 gpurun_out/d_trace_gemm_f16_fc2.log
