#!/bin/bash
# Round-2 visit T: final validation of the committed tree (fused projections, packed Sinkhorn sweep, training step).
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/t_pytest_gpu.log
tail -3 gpurun_out/t_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; tail -1 gpurun_out/t_smoke.log
timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t_bench_ref.json 2> gpurun_out/t_bench_ref.err
for wl in C1 C2 C5; do
  timeout 600 python bench.py --no-cpu-baseline --steps 8 --workload $wl > gpurun_out/t_bench_$wl.json 2> gpurun_out/t_bench_$wl.err
done
python - <<'PY'
import json
for n in ('t_bench','t_bench_ref','t_bench_C1','t_bench_C2','t_bench_C5'):
    try:
        d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
        r=d.get('roofline') or {}; rs=d.get('roofline_sinkhorn') or {}
        print(n, round(d['value'],2), 'pairs/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), '; launches', d.get('gpu_launches'), '; attn', r.get('ms_per_launch'), r.get('achieved'), r.get('traffic'), '; sink', rs.get('frac'), rs.get('ms_per_launch'), (d.get('verified') or {}).get('matches0_identical'), d.get('clocks'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-400:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify --cuda-graph 0 > gpurun_out/t_ncu_bench.log 2>&1
python scripts/agg_launches.py gpurun_out/t_launches.csv > gpurun_out/t_launches_agg.txt 2>&1; head -16 gpurun_out/t_launches_agg.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:"sinkhorn_kernel|linear_f16_kernel" -s 40 -c 12 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify --cuda-graph 0 2>&1 | grep -E "^  [a-z_ ]*(sinkhorn_kernel|linear_f16_kernel)|gpu__time|dram__|pipe_tensor|issue_active|pipe_xu|registers|warps_active" > gpurun_out/t_ncu_kernels.txt
head -30 gpurun_out/t_ncu_kernels.txt
timeout 300 python scripts/train_step_time.py 8 1024 1024 9 20 2>&1 | tail -2 | tee gpurun_out/t_train_step.txt
