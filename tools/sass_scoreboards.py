#!/usr/bin/env python
"""For every loop of a cuobjdump -sass listing that contains 128-bit global loads: where the loads sit in the loop body, which
scoreboards they write (decoded from the control bits of the instruction words) and where the first waits on those
scoreboards are -- shows whether a register software pipeline survived ptxas.

    cuobjdump -sass -fun <mangled kernel> slotdiffusion_b200/build/slot_attention_resident.o > k.sass; python tools/sass_scoreboards.py k.sass
"""
import re,sys
lines=open(sys.argv[1]).read().split('\n')
ins=[]
i=0
while i<len(lines):
    m=re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/',lines[i])
    if m and i+1<len(lines):
        m2=re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/',lines[i+1])
        if m2:
            hi=int(m2.group(1),16)
            ctrl=(hi>>41)&0x7fffff
            ins.append((int(m.group(1),16),m.group(2).strip(),ctrl&0xf,(ctrl>>5)&7,(ctrl>>8)&7,(ctrl>>11)&0x3f))
            i+=2; continue
    i+=1
# summarize each loop containing LDG.E.128: find backward branches
for k,(a,t,st,wb,rb,wt) in enumerate(ins):
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a:
            body=[x for x in ins if tgt<=x[0]<=a]
            nl=sum('LDG.E.128' in x[1] for x in body)
            if nl:
                nf=sum(x[1].startswith('FFMA') for x in body)
                pos=[ (x[0]-tgt)//16 for x in body if 'LDG.E.128' in x[1]]
                sbs=sorted(set(x[3] for x in body if 'LDG.E.128' in x[1]))
                firstwait=[ (x[0]-tgt)//16 for x in body if any((x[5]>>s)&1 for s in sbs)][:3]
                print('loop %05x-%05x: %d instr, %d FFMA, %d LDG at instr %s, SBs %s, first waits on them at %s'%(tgt,a,len(body),nf,nl,pos if len(pos)<=12 else (pos[:6],'...',pos[-3:]),sbs,firstwait))
