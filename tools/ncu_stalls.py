#!/usr/bin/env python
"""Summarise the source page of an ncu report: top stall sites and sample / instruction counts per SASS region.
    python tools/ncu_stalls.py gpurun_out/ncu_x.ncu-rep [bucket]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, '# Samples') for r in data)
print('kernel', rows[0][1][:100])
print('total samples', tot)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:25]:
    st = sorted(((f(r, c), c) for c in stall_cols), reverse=True)[:2]
    print(int(f(r, '# Samples')), r[ix['Address']][-5:], r[ix['Source']][:80], st)
acc = accI = 0
start = None
for n, r in enumerate(data):
    if start is None:
        start = r[ix['Address']][-5:]
    acc += f(r, '# Samples')
    accI += f(r, 'Instructions Executed')
    if (n + 1) % bucket == 0 or n == len(data) - 1:
        if acc or accI:
            print(start, r[ix['Address']][-5:], 'samples', int(acc), 'instr', int(accI), '|', data[max(0, n - bucket // 2)][ix['Source']][:60])
        acc = accI = 0
        start = None
