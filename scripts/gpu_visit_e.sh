#!/bin/bash
# programmatic dependent launch (OG_PDL=1): parity + timing against OG_PDL=0, with and without CUDA-graph replay
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
OG_PDL=1 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/e_tests_pdl.log 2>&1; echo "rc=$?" >> gpurun_out/e_tests_pdl.log
tail -3 gpurun_out/e_tests_pdl.log
for cfg in 0,1 1,1 0,0 1,0; do
  pdl=${cfg%,*}; g=${cfg#*,}
  OG_PDL=$pdl timeout 300 python bench.py --steps 8 --no-cpu-baseline --cuda-graph $g > gpurun_out/e_bench_pdl${pdl}_g${g}.json 2> gpurun_out/e_bench_pdl${pdl}_g${g}.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/e_bench_pdl${pdl}_g${g}.json'))
    print('pdl=$pdl graph=$g', round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'])
except Exception as e:
    print('pdl=$pdl graph=$g failed', e)
PY
done
OG_PDL=1 timeout 300 python bench.py --workload C1 --steps 20 --no-cpu-baseline > gpurun_out/e_bench_C1_pdl1.json 2>/dev/null
OG_PDL=0 timeout 300 python bench.py --workload C1 --steps 20 --no-cpu-baseline > gpurun_out/e_bench_C1_pdl0.json 2>/dev/null
python - <<PY
import json
for p in (0, 1):
    try:
        d = json.load(open('gpurun_out/e_bench_C1_pdl%d.json' % p)); print('C1 pdl=%d' % p, round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms')
    except Exception as e:
        print('C1 failed', p, e)
PY
