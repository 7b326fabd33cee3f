#!/bin/bash
set +e
mkdir -p gpurun_out
OG_ATTN_PAIR=1 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 300 -k "attention_tc" 2>&1 | tail -12 > gpurun_out/pytest_pair.log
tail -6 gpurun_out/pytest_pair.log
OG_ATTN_PAIR=1 timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_all.log
OG_ATTN_PAIR=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err
OG_ATTN_PAIR=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopair.json 2> gpurun_out/bench_nopair.err
python - <<'PY'
import json
for n in ('pair','nopair'):
    try:
        d=json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms; attn', round(d['roofline']['ms_per_launch'],4), 'ms', round(d['roofline']['achieved'],1), 'TF/s')
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/bench_{n}.err').read()[-600:])
PY
