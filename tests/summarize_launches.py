"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name']
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = name.split('(')[0]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6, 'ns ': 1}.get(unit, 1)
    tot[name] += ns; cnt[name] += 1
T = sum(tot.values())
print(f'total {T/1e6:.3f} ms over {sum(cnt.values())} launches')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f'{v/1e6:9.3f} ms {100*v/T:5.1f}%  n={cnt[k]:5d}  avg={v/cnt[k]/1e3:8.1f} us  {k[:90]}')
