#!/bin/bash
# Round-2 visit X (final, ~10 GPU-minutes left): visit W found the early S hand-back SLOWER (0.344 vs 0.336 ms) and the tile-end role
# swap (M) the only addition that gained on top of it; here M / P (pair barriers) are measured WITHOUT the early hand-back, the best
# verified build gets the full GPU suite, the headline bench and the other BASELINE configurations.
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
O=gpurun_out
AB=openglue_b200/ab
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
: > $O/x_ab_attention.jsonl
for v in default M P MP default; do
  if [ "$v" = default ]; then L=openglue_b200/libopenglue_b200.so; else L=$AB/lib_$v.so; fi
  OG_LIB=$L timeout 150 python scripts/ab_attention.py 2> $O/x_ab_$v.err | tail -1 >> $O/x_ab_attention.jsonl
  el "ab_attention $v done"
done
python - <<'P'
import json
for line in open('gpurun_out/x_ab_attention.jsonl'):
    try: d = json.loads(line)
    except Exception: print('unparsable:', line[:200]); continue
    print(d['lib'], d.get('ok'), round(d.get('ms_self_32x4x2048x2048', 0), 4), round(d.get('ms_cross_16x4x2048x2048', 0), 4), max(d['parity'].values()) if d.get('parity') else None, d.get('error'))
P
BEST=$(python - <<'P'
import json
res, base = {}, []
for line in open('gpurun_out/x_ab_attention.jsonl'):
    try: d = json.loads(line)
    except Exception: continue
    name = 'default' if d['lib'].endswith('libopenglue_b200.so') else d['lib'].split('lib_')[-1][:-3]
    if d.get('ok') and 'ms_self_32x4x2048x2048' in d:
        t = d['ms_self_32x4x2048x2048'] + d['ms_cross_16x4x2048x2048']
        if name == 'default': base.append(t)
        else: res[name] = t
best = 'default'
if base and res:
    b = min(res, key=res.get)
    if res[b] < min(base) * 0.99: best = b          # the default was measured twice (first and last): its spread is the noise floor
print(best)
P
)
el "best attention build: $BEST"
if [ "$BEST" = default ]; then CHOSEN=openglue_b200/libopenglue_b200.so; else CHOSEN=$AB/lib_$BEST.so; fi
echo "$BEST $CHOSEN" > $O/x_chosen.txt
OG_LIB=$CHOSEN timeout 400 python -m pytest tests -m gpu -q --timeout 300 -rf 2>&1 | tail -30 > $O/x_pytest_gpu.log; tail -4 $O/x_pytest_gpu.log; el "pytest"
OG_LIB=$CHOSEN timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/x_bench.json 2> $O/x_bench.err; el "bench"
OG_LIB=$CHOSEN timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/x_smoke.log; el "smoke"
for wl in C1 C2 C5; do
  OG_LIB=$CHOSEN timeout 120 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > $O/x_bench_$wl.json 2> $O/x_bench_$wl.err; el "bench $wl"
done
OG_LIB=$CHOSEN timeout 240 ncu --set full --clock-control none --import-source on -k regex:attention_f16t -s 16 -c 1 -o $O/x_prof_attn -f python scripts/ab_attention.py > $O/x_prof_attn.log 2>&1; el "ncu full attention"
python - <<'P'
import json
for n in ('x_bench', 'x_bench_C1', 'x_bench_C2', 'x_bench_C5'):
    try:
        d = json.load(open(f'gpurun_out/{n}.json'))
        print(n, round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms; e2e', round(d['e2e']['value'], 1), '; attn ms', round(d['roofline']['ms_per_launch'], 4),
              'frac', round(d['roofline']['frac'], 4), '; sink ms', round(d['roofline_sinkhorn']['ms_per_launch'], 3), 'frac', round(d['roofline_sinkhorn']['frac'], 3),
              (d.get('verified') or {}).get('matches0_identical'), d['clocks'])
    except Exception as e:
        print(n, 'failed', e)
        try: print(open(f'gpurun_out/{n}.err').read()[-400:])
        except Exception: pass
P
el "end"
