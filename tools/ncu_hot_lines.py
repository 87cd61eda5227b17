#!/usr/bin/env python
"""SASS lines of the first kernel of an `ncu --page source --csv --print-source sass` dump with at least N executions.

    python tools/ncu_hot_lines.py src.csv 300000     (index, executed, stall samples, instruction)"""
import csv,sys
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
lo=int(sys.argv[2])
for i,r in enumerate(data):
    n=int(r[ia])
    if n>=lo: print(f"{i:5d} {n:8d} {int(r[ist]):5d}  {r[isrc].strip()}")
