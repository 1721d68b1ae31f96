"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per kernel section, top instructions by stall samples and
per-opcode totals.   python tools/ncu_src_summary.py file.csv [top_n]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sections, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        sections.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
for s in sections:
    idx = {h: i for i, h in enumerate(s['hdr'])}
    data = s['data']
    tot = sum(int(r[idx['# Samples']]) for r in data)
    print('=====', s['name'], 'samples', tot, 'instructions', len(data))
    keys = [k for k in s['hdr'] if k.startswith('stall_') and 'Not Issued' not in k]
    agg = Counter()
    for r in data:
        for k in keys:
            agg[k] += int(r[idx[k]])
    print('stall totals:', {k.replace('stall_', ''): v for k, v in agg.most_common(10)})
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:top_n]:
        st = {k.replace('stall_', ''): int(r[idx[k]]) for k in keys if int(r[idx[k]])}
        print(f"{r[idx['# Samples']]:>5} {r[idx['Source']].strip()[:64]:<64} exec {r[idx['Instructions Executed']]:>7} {st}")
    c, ex = Counter(), Counter()
    for r in data:
        toks = r[idx['Source']].split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        c[op] += int(r[idx['# Samples']]); ex[op] += int(r[idx['Instructions Executed']])
    print('samples by opcode:', c.most_common(12))
    print('executed by opcode:', ex.most_common(12))
