#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that show what hardware path a kernel uses (tcgen05 MMA = UTCHMMA / UTCQMMA,
TMEM loads = LDTM, TMA tensor loads = UTMALDG, bulk copies = UBLKCP, cluster barriers = UCGABAR, DSMEM = *.CLUSTER ...).

    python tools/sass_summary.py [slotdiffusion_b200/libsdb200.so] > profiles/<round>_sass_summary.md
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'slotdiffusion_b200/libsdb200.so'
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'UCGABAR', 'MAPA', 'FFMA', 'HMMA']
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.split('\n'):
    m = re.match(r'\s+Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        counts[cur]['_n'] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
        if '.2CTA' in op:
            counts[cur]['2CTA'] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.split('\n')
print('| kernel | instr | ' + ' | '.join(KEYS + ['.2CTA']) + ' |')
print('|---|---:|' + '---:|' * (len(KEYS) + 1))
rows = []
for name, d in zip(demangle, counts.values()):
    short = re.sub(r'\(.*', '', name).replace('void ', '')
    if not any(d[k] for k in KEYS if k != 'FFMA'):
        continue
    rows.append((short, d))
for short, d in sorted(rows, key=lambda r: r[0]):
    print('| `%s` | %d | ' % (short[:90], d['_n']) + ' | '.join(str(d[k]) if d[k] else '' for k in KEYS + ['2CTA']) + ' |')
