"""Per-kernel SASS opcode histogram of libinfgen_b200.so (evidence that the tensor-core / TMA instructions are there).

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.md

Mnemonics (B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk,
UTMALDG = TMA tensor load, LDGSTS = cp.async, SYNCS = mbarrier, HMMA = legacy mma.sync, FFMA = fp32 FMA."""
import os, re, subprocess, sys
from collections import Counter, OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'infgen_b200', 'lib', 'libinfgen_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
kern, cur = OrderedDict(), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('infgen::', '')
        cur = kern.setdefault(name, Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'LDGSTS', 'SYNCS', 'HMMA', 'FFMA', 'FMUL', 'FADD', 'MUFU',
       'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'BAR', 'LDL', 'STL']
print('# SASS opcode histogram per kernel (`cuobjdump -sass infgen_b200/lib/libinfgen_b200.so`, sm_100a)\n')
print('| kernel | instr | ' + ' | '.join(KEY) + ' |')
print('|---|---|' + '---|' * len(KEY))
for name, c in sorted(kern.items(), key=lambda kv: -sum(kv[1].values())):
    tot = sum(c.values())
    print(f'| `{name}` | {tot} | ' + ' | '.join(str(sum(v for k, v in c.items() if k.startswith(key))) for key in KEY) + ' |')
