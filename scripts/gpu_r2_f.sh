#!/bin/bash
# Round-2 visit F: fp16 attention forms (2 = ping-pong warpgroups, 1 = two teams of two warpgroups, 0 = one team) after the
# setmaxnreg / function-call fix: parity, timing, traces.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -s --timeout 500 2>&1 | grep -v "^$" | tail -40 > gpurun_out/f_$name.log; echo "== $name"; grep -E "^\[|passed|failed|rror" gpurun_out/f_$name.log | tail -12; }
OG_ATTN_FORM=2 run attn2 tests/test_gpu_f16.py -k "attention_f16"
OG_ATTN_FORM=1 run attn1 tests/test_gpu_f16.py -k "attention_f16"
OG_ATTN_FORM=2 run path2 tests/test_gpu_parity.py -k "(reference_big and fp16x3) or (matches_oracle and fp16x3)"
for form in 2 1 0; do
  OG_ATTN_FORM=$form timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline --steps 6 > gpurun_out/f_bench_f16_form$form.json 2> gpurun_out/f_bench_f16_form$form.err
done
python - <<'PY'
import json
for n in ('f_bench_f16_form2','f_bench_f16_form1','f_bench_f16_form0'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms; attn', round(d['roofline']['ms_per_launch'],4), 'ms', round(d['roofline']['achieved'],1), 'TF/s; verified', d['verified']['matches0_identical'])
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-400:])
PY
OG_ATTN_FORM=2 timeout 200 python scripts/trace_f16.py attn > gpurun_out/f_trace_attn_f16p.log 2>&1; head -26 gpurun_out/f_trace_attn_f16p.log | cut -c1-200
OG_ATTN_FORM=1 timeout 200 python scripts/trace_f16.py attn > gpurun_out/f_trace_attn_f16t.log 2>&1; head -16 gpurun_out/f_trace_attn_f16t.log | cut -c1-200
