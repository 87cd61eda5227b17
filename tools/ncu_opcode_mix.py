#!/usr/bin/env python
"""Opcode mix of the first kernel of an `ncu --page source --csv --print-source sass` dump, weighted by executed count.

    ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_opcode_mix.py src.csv"""
import csv,sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; data=[]
k=0
for r in rows:
    if r and r[0]=="Kernel Name":
        k+=1
        if k>1: break
        continue
    if r and r[0]=="Address": hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(r)
ia=hdr.index("Instructions Executed"); isrc=hdr.index("Source"); ist=hdr.index("Warp Stall Sampling (All Samples)")
tot=sum(int(r[ia]) for r in data); tots=sum(int(r[ist]) for r in data)
print("instrs",len(data),"executed",tot,"samples",tots)
c=Counter(int(r[ia]) for r in data)
print(c.most_common(8))
h=Counter(); hs=Counter()
for r in data:
    op=r[isrc].split()
    op=[o for o in op if not o.startswith('@')][0].split('.')[0]
    h[op]+=int(r[ia]); hs[op]+=int(r[ist])
for op,n in h.most_common(30): print(f"{op:12s} {n/tot*100:5.1f} % of executed   {hs[op]/tots*100:5.1f} % of stall samples")
