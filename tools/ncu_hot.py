#!/usr/bin/env python
"""Hot instructions of a kernel from `ncu --page source --csv` (SASS view): top stall-sample addresses."""
import csv, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
iS = hdr.index('# Samples'); iSrc = hdr.index('Source'); iEx = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS] or 0) for r in data)
print('total samples', tot, ' instructions', len(data))
idx = {id(r): k for k, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print(f"{idx[id(r)]:5d} {int(r[iS]):7d} {100*int(r[iS])/tot:5.1f}%  ex={r[iEx]:>9s}  {r[iSrc].strip()[:70]:70s} {st}")
