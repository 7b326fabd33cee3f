#!/bin/bash
# Round-2 multi-GPU visit (gpurun --gpus 2): BASELINE configs[3] as specified (32 pairs / GPU, labels -> forward -> reference
# criterion per rank, NCCL mean of the loss on a side stream) and the headline workload on 2 GPUs.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m_gpus.txt
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload C4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_C4_n$N.json 2> gpurun_out/m_bench_C4_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_C3_n$N.json 2> gpurun_out/m_bench_C3_n$N.err
timeout 600 python bench.py --gpus 1 --workload C4 --steps 6 --no-cpu-baseline > gpurun_out/m_bench_C4_n1.json 2> gpurun_out/m_bench_C4_n1.err
python - <<PY
import json
for n in ('m_bench_C4_n$N','m_bench_C3_n$N','m_bench_C4_n1'):
    try:
        d=[json.loads(l) for l in open(f'gpurun_out/{n}.json').read().strip().splitlines() if l.startswith('{')][-1]
        print(n, round(d['value'],1), 'pairs/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'n_gpus', d['n_gpus'], 'loss', d.get('loss'), 'verified', (d.get('verified') or {}).get('matches0_identical'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/{n}.err').read()[-600:])
PY
