#!/usr/bin/env python
"""Segment an `ncu --page source --csv` SASS listing at barriers/atomics to see where samples and instructions go."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; data=[r for r in rows[hi+1:] if len(r)>10 and r[0].startswith('0x')]
iS=hdr.index('# Samples'); iE=hdr.index('Instructions Executed'); iSrc=hdr.index('Source'); iT=hdr.index('Avg. Threads Executed')
tot_s=sum(int(r[iS]) for r in data); tot_e=sum(int(r[iE]) for r in data)
print('total samples',tot_s,'total inst',tot_e,'n sass',len(data))
marks=tuple(sys.argv[2].split(',')) if len(sys.argv)>2 else ('BAR','MATCH','EXIT','ATOMG','RED','ATOMS','REDUX','BRA')
seg_s=seg_e=0
for n,r in enumerate(data):
    s=int(r[iS]); e=int(r[iE]); seg_s+=s; seg_e+=e
    src=r[iSrc].strip(); t=src.split()
    op=t[1] if t and t[0].startswith('@') and len(t)>1 else (t[0] if t else '')
    if op.startswith(marks) and (100*seg_s/tot_s>0.8 or 100*seg_e/tot_e>0.8 or not op.startswith('BRA')):
        print('%5d %-46s exec=%9d thr=%5s | seg samples %5.1f%% inst %5.1f%%'%(n, src[:46], e, r[iT], 100*seg_s/tot_s, 100*seg_e/tot_e)); seg_s=seg_e=0
print('tail seg samples %.1f%% inst %.1f%%'%(100*seg_s/tot_s, 100*seg_e/tot_e))
