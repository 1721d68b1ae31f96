"""DRAM / L2 traffic and duration per captured kernel from `ncu -i X.ncu-rep --page raw --csv` files ->
profiles/r2_ncu_traffic.json (read by bench.py for `roofline.traffic`).

    python tools/ncu_traffic.py out.json raw1.csv [raw2.csv ...]"""
import csv, json, re, sys
from collections import defaultdict


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return 0.0


SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'msecond': 1e3, 'usecond': 1.0,
         'nsecond': 1e-3, 'second': 1e6}
acc = defaultdict(lambda: defaultdict(float))
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('infgen::', '').replace('void ', '')
        a = acc[name]
        a['launches'] += 1
        for key, out in (('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
                         ('lts__t_bytes.sum', 'l2_bytes'), ('gpu__time_duration.sum', 'duration_us')):
            if key in ix:
                a[out] += num(r[ix[key]]) * SCALE.get(units[ix[key]], 1.0)
        a['source'] = path.split('/')[-1]
res = {'how': 'ncu --set full --clock-control none, per launch averages over the captured launches (cold-cache, serialised)',
       'kernels': {}}
for name, a in acc.items():
    n = a['launches']
    res['kernels'][name] = {'launches_captured': int(n), 'dram_bytes_per_launch': (a['dram_read'] + a['dram_write']) / n,
                            'dram_read_per_launch': a['dram_read'] / n, 'dram_write_per_launch': a['dram_write'] / n,
                            'l2_bytes_per_launch': a['l2_bytes'] / n, 'duration_us_per_launch': a['duration_us'] / n,
                            'source': a['source']}
json.dump(res, open(sys.argv[1], 'w'), indent=1)
print(json.dumps(res, indent=1))
