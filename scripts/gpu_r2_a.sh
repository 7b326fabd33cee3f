#!/bin/bash
# Round-2 visit A: parity at the quoted configs (new fixtures), headline bench with self-verification, reference arm,
# launch list + ncu captures of the cross-attention launch and of the Q/K/V projection GEMM variants.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_nvsmi.txt 2>&1
lscpu | head -20 > gpurun_out/a_cpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/a_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_tc|linear_tc2" -c 8 -o gpurun_out/a_prof_evidence -f python scripts/prof_ops.py evidence 1 > gpurun_out/a_prof_evidence.log 2>&1
tail -5 gpurun_out/a_pytest_gpu.log; tail -2 gpurun_out/a_smoke.log
python - <<'PY'
import json
for n in ('a_bench','a_bench_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],3), d['unit'], round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['value'],2), d.get('verified'), d.get('clocks'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-800:])
PY
