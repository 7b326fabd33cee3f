#!/bin/bash
# Full validation visit: all GPU tests, smoke, headline bench (+cpu baseline), reference arm, other configs.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for wl in C1 C2 C5; do
  timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
done
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2
python - <<'PY'
import json
for n in ('bench','bench_ref','bench_C1','bench_C2','bench_C5'):
    try:
        d=json.load(open(f'gpurun_out/{n}.json'))
        print(n, round(d['value'],2), d['unit'], round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['value'],2), d.get('cpu_baseline',{}).get('value'), d.get('clocks'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-500:])
PY
