#!/bin/bash
# Round-2 visit C: new Sinkhorn kernel (W warps per row, 2 CTAs/SM), fp16x3 launch list, event traces of the fp16 kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -s --timeout 500 2>&1 | grep -v "^$" | tail -40 > gpurun_out/c_$name.log; echo "== $name"; grep -E "^\[|passed|failed|error|Error" gpurun_out/c_$name.log | tail -12; }
run sinkhorn tests/test_gpu_parity.py -k "sinkhorn_operator or headline_shape"
run path tests/test_gpu_parity.py -k "(reference_big and fp16x3-1) or matches_oracle or golden"
for occ in 2 1; do
  OG_SINK_OCC=$occ timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 6 > gpurun_out/c_bench_f16_occ$occ.json 2> gpurun_out/c_bench_f16_occ$occ.err
  OG_SINK_OCC=$occ timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 6 --workload C2 > gpurun_out/c_bench_f16_C2_occ$occ.json 2> gpurun_out/c_bench_f16_C2_occ$occ.err
done
python - <<'PY'
import json
for n in ('c_bench_f16_occ2','c_bench_f16_occ1','c_bench_f16_C2_occ2','c_bench_f16_C2_occ1'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms; attn', round(d['roofline']['ms_per_launch'],4), 'ms; sinkhorn', round(d['roofline_sinkhorn']['ms_per_launch'],3), 'ms', round(d['roofline_sinkhorn']['frac'],3), d.get('verified',{}) and d['verified'].get('matches0_identical'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-600:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/c_launches_f16.csv python bench.py --precision fp16x3 --steps 1 --warmup 1 --no-cpu-baseline --no-verify --cuda-graph 0 > gpurun_out/c_ncu_bench.log 2>&1
python scripts/agg_launches.py gpurun_out/c_launches_f16.csv > gpurun_out/c_launches_f16_agg.txt 2>&1; head -30 gpurun_out/c_launches_f16_agg.txt
timeout 200 python scripts/trace_f16.py attn > gpurun_out/c_trace_attn_f16.log 2>&1; head -24 gpurun_out/c_trace_attn_f16.log
timeout 200 python scripts/trace_f16.py gemm fc2 > gpurun_out/c_trace_gemm_f16_fc2.log 2>&1; tail -30 gpurun_out/c_trace_gemm_f16_fc2.log
timeout 200 python scripts/trace_f16.py gemm q > gpurun_out/c_trace_gemm_f16_q.log 2>&1; tail -12 gpurun_out/c_trace_gemm_f16_q.log
