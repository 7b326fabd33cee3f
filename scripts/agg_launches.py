"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(row['Metric Unit'], 1e-6)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f'{"kernel":58s} {"n":>5s} {"total ms":>10s} {"share":>7s} {"avg ms":>9s}')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:58]:58s} {v[0]:5d} {v[1]:10.3f} {v[1]/tot*100:6.1f}% {v[1]/v[0]:9.4f}')
print(f'{"total":58s} {sum(v[0] for v in agg.values()):5d} {tot:10.3f}')
