#!/bin/bash
set +e
mkdir -p gpurun_out; rm -f gpurun_out/tc_probe_pair.log
for pm in 1 0; do
  echo "== OG_GEMM_PAIR=$pm" >> gpurun_out/tc_probe_pair.log
  OG_GEMM_PAIR=$pm timeout 600 python scripts/tc_probe.py v2_k256 v2_tails_concat v2_resid_split_t v2_batchedB v2_qkv_shape_time v2_fc1_shape_time v2_fc2_shape_time v2_score_shape_time >> gpurun_out/tc_probe_pair.log 2>&1
done
cat gpurun_out/tc_probe_pair.log
OG_GEMM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_all.log
OG_GEMM_PAIR=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gpair.json 2> gpurun_out/bench_gpair.err
OG_GEMM_PAIR=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gnopair.json 2> gpurun_out/bench_gnopair.err
python - <<'PY'
import json
for n in ('gpair','gnopair'):
    try:
        d=json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms')
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/bench_{n}.err').read()[-600:])
PY
