#!/bin/bash
# Round-2 visit Y: the driver's own sequence on the final default build - smoke, `bench.py` with its defaults (CPU baseline leg on the
# staged reference, oracle/_ref) and the reference arm.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/y_smoke.log; el smoke
timeout 300 python bench.py > $O/y_bench.json 2> $O/y_bench.err; el bench
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/y_bench_ref.json 2> $O/y_bench_ref.err; el "reference arm"
python - <<'P'
import json
for n in ('y_bench', 'y_bench_ref'):
    try:
        d = json.load(open(f'gpurun_out/{n}.json'))
        cb = d.get('cpu_baseline', {})
        print(n, round(d['value'], 3), d['unit'], round(d['ms_per_step'], 3), 'ms/step; e2e', round(d['e2e']['value'], 3), '; cpu_baseline', cb.get('kind'), cb.get('value'), cb.get('cores'),
              '; roofline', (d.get('roofline') or {}).get('frac'), (d.get('verified') or {}).get('matches0_identical'), d.get('clocks'))
        print('   sample:', cb.get('sample', '')[:300])
    except Exception as e:
        print(n, 'failed', e)
        try: print(open(f'gpurun_out/{n}.err').read()[-600:])
        except Exception: pass
P
el end
