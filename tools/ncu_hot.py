#!/usr/bin/env python
"""Per-instruction stall samples of one kernel in an .ncu-rep (source page, SASS): the hottest instructions in program order, and the
cumulative share of samples at chosen marker instructions (barriers, MMA, tensor-memory loads, MUFU, shuffles).

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [n_top]
"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
iS, iSrc = hdr.index("# Samples"), hdr.index("Source")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {n: hdr.index(n) for n in names}
tot = sum(int(r[iS] or 0) for r in data)
print(rows[0][1], "total samples", tot)
cum = [0]
for r in data:
    cum.append(cum[-1] + int(r[iS] or 0))
print("-- markers (cumulative share of samples BEFORE the instruction)")
last = None
for i, r in enumerate(data):
    s = r[iSrc].strip()
    m = re.search(r"(SYNCS\.\w+|BAR\.\w+|UTCHMMA|UTCBAR|LDTM|MUFU\.\w+|SHFL|LDGSTS|LDGDEPBAR|EXIT|FENCE|STS\.128|LDS)", s)
    if not m:
        continue
    kind = m.group(1)
    if kind != last:
        print(f"{i:5d} cum={100 * cum[i] / tot:5.1f}%  {s[:80]}")
        last = kind
print("-- hottest instructions")
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS] or 0))[:n_top]
for i in sorted(top):
    r = data[i]
    st = sorted(((n, int(r[idx[n]] or 0)) for n in names), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {100 * int(r[iS]) / tot:5.1f}%  {r[iSrc].strip()[:70]:70s} {st}")
