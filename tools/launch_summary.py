"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
    python tools/launch_summary.py launches.csv > profiles/rN_launch_summary.csv"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('infgen::', '')
    v = float(r[ix['Metric Value']].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ix['Metric Unit']], 1e-3)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print('kernel,launches,total_us,avg_us,share')
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f'{k},{cnt[k]},{tot[k]:.1f},{tot[k] / cnt[k]:.2f},{tot[k] / s:.4f}')
print(f'TOTAL,{sum(cnt.values())},{s:.1f},,1.0')
