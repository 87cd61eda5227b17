#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        d[r[ki]][0] += 1
        d[r[ki]][1] += float(r[vi].replace(",", ""))
    mine = ("corr_", "fi_forward", "fi_backward", "projection_", "interp_", "sepconv")
    ours = {k: v for k, v in d.items() if any(m in k for m in mine)}
    tot = sum(v[1] for v in ours.values())
    print(f"| kernel (libvfidkr_b200.so) | launches | total us | us / launch | share of our kernels |")
    print("|---|---|---|---|---|")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        name = k.split("(")[0].replace("void ", "")
        print(f"| `{name}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot * 100:.1f} % |")
    other = sum(v[1] for k, v in d.items() if k not in ours)
    print(f"\nother kernels in the capture (torch input generation etc.): {other / 1e3:.1f} us in {sum(v[0] for k, v in d.items() if k not in ours)} launches")


if __name__ == "__main__":
    main(sys.argv[1])
