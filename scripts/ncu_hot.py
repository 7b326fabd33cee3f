"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu_hot.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: s = int(r[isamp])
    except ValueError: continue
    st = {hdr[i]: int(r[i] or 0) for i in stall_cols}
    data.append((s, r[isrc].strip(), int(r[iexec] or 0), st, len(data)))
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for d in data:
    for k, v in d[3].items(): agg[k] = agg.get(k, 0) + v
print('stall mix:', ', '.join(f'{k[6:]}={v*100/tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 / tot > 1))
for s, src, ex, st, idx in sorted(data, key=lambda d: -d[0])[:top]:
    main = max(st.items(), key=lambda kv: kv[1])
    print(f'{idx:5d} {s*100/tot:5.1f}% exec={ex:9d}  {main[0][6:]:12s} {src[:90]}')
