#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "cuda_graph or host_buffers" 2>&1 | tail -12
for wl in C1 C3; do for g in 0 1; do
  timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline --cuda-graph $g > gpurun_out/bench_${wl}_g$g.json 2> gpurun_out/bench_${wl}_g$g.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${wl}_g$g.json')); print('$wl graph=$g', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1))
except Exception as e:
    print('$wl graph=$g failed'); print(open('gpurun_out/bench_${wl}_g$g.err').read()[-800:])
PY
done; done
