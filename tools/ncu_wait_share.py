#!/usr/bin/env python
"""Share of executed warp instructions that sit in mbarrier wait loops, from the source page of an ncu report.
    python tools/ncu_wait_share.py gpurun_out/full_r1d.ncu-rep [kernel-name regex]
A wait loop is recognised structurally: the backward-branch region around every SYNCS.PHASECHK (mbarrier try_wait /
test_wait) whose instructions execute more often than the busiest non-loop line of the kernel."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else 'slot_attend_fused'
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr_at = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[hdr_at[0]]
body = rows[hdr_at[0] + 1:(hdr_at[1] - 1 if len(hdr_at) > 1 else len(rows))]
ie, src = hdr.index('Instructions Executed'), hdr.index('Source')
n = [int(r[ie]) if r[ie].isdigit() else 0 for r in body]
total = sum(n)
waits = [i for i, r in enumerate(body) if 'SYNCS.PHASECHK' in r[src]]
in_loop = set()
for w in waits:
    # contiguous neighbourhood executed at least half as often as the poll itself = the poll loop body
    for step in (-1, 1):
        i = w
        while 0 <= i < len(body) and n[i] * 2 >= n[w] and n[w] > 0 and abs(i - w) < 40:
            in_loop.add(i)
            i += step
loop_instr = sum(n[i] for i in in_loop)
print(rows[0][1][:90])
print(f'warp instructions executed: {total}; in mbarrier poll loops: {loop_instr} ({100 * loop_instr / total:.1f} %)')
print('poll sites (executions of the wait instruction, loop-body instructions per poll):')
for w in waits:
    if n[w] < 1000:
        continue
    lo = w
    while lo - 1 in in_loop and abs(lo - 1 - w) < 40:
        lo -= 1
    hi = w
    while hi + 1 in in_loop and abs(hi + 1 - w) < 40:
        hi += 1
    print(f'  {n[w]:>8}  x ~{sum(n[lo:hi + 1]) / max(n[w], 1):4.1f}   {re.sub(" +", " ", body[w][src].strip())[:70]}')
