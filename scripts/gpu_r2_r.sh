#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_f16.py -x -q -m gpu 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for fz in 1 0; do
  OG_FUSE_QKV=$fz timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r_bench_fuse$fz.json 2> gpurun_out/r_bench_fuse$fz.err
  OG_FUSE_QKV=$fz timeout 300 python bench.py --workload C1 --steps 30 --warmup 5 > gpurun_out/r_bench_C1_fuse$fz.json 2> gpurun_out/r_bench_C1_fuse$fz.err
done
python - <<'P'
import json
for f in ('r_bench_fuse1','r_bench_fuse0','r_bench_C1_fuse1','r_bench_C1_fuse0'):
    try:
        d=[json.loads(l) for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1]
        print(f, d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
P
