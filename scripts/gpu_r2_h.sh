#!/bin/bash
# Round-2 visit H: full validation with the fp16x3 defaults + evidence captures.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/h_pytest_gpu.log
tail -4 gpurun_out/h_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/h_smoke.log 2>&1; tail -1 gpurun_out/h_smoke.log
timeout 900 python bench.py > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err
for wl in C1 C2 C5; do
  timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload $wl > gpurun_out/h_bench_$wl.json 2> gpurun_out/h_bench_$wl.err
done
timeout 600 python bench.py --no-cpu-baseline --steps 8 --precision tf32x3 > gpurun_out/h_bench_tf32x3.json 2> gpurun_out/h_bench_tf32x3.err
python - <<'PY'
import json
for n in ('h_bench','h_bench_ref','h_bench_C1','h_bench_C2','h_bench_C5','h_bench_tf32x3'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        r=d.get('roofline') or {}; rs=d.get('roofline_sinkhorn') or {}
        print(n, round(d['value'],2), 'pairs/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), '; attn', r.get('ms_per_launch'), r.get('achieved'), '; sink', rs.get('frac'), (d.get('verified') or {}).get('matches0_identical'), d.get('clocks'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-400:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify --cuda-graph 0 > gpurun_out/h_ncu_bench.log 2>&1
python scripts/agg_launches.py gpurun_out/h_launches.csv > gpurun_out/h_launches_agg.txt 2>&1; head -14 gpurun_out/h_launches_agg.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_f16|linear_f16" -c 9 -o gpurun_out/h_prof_f16 -f python scripts/prof_ops.py f16 1 > gpurun_out/h_prof_f16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_kernel -s 1 -c 1 -o gpurun_out/h_prof_sinkhorn -f python scripts/prof_ops.py sinkhorn 2 > gpurun_out/h_prof_sinkhorn.log 2>&1
ls -la gpurun_out/h_prof*.ncu-rep
