#!/usr/bin/env python
"""Aggregate the per-instruction warp-stall samples of an .ncu-rep (source page, SASS view) into blocks of BLK
instructions: share of samples, executed warp instructions, dominant opcodes and stall reasons.

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep [BLK]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# the source page lists one block per profiled launch: ["Kernel Name", name] / header / instructions ...
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    hdr, data = rows[a + 1], [r for r in rows[a + 2:b] if len(r) == len(rows[a + 1])]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in data)
    print(rows[a][1]); print("total samples", tot, "SASS instructions", len(data), "warp instructions executed", sum(int(r[ii]) for r in data))
    for b0 in range(0, len(data), blk):
        chunk = data[b0:b0 + blk]
        s_ = sum(int(r[si]) for r in chunk)
        if s_ < tot * 0.01:
            continue
        ops = collections.Counter(re.sub(r"^\s*(@!?U?P\d+\s+)?", "", r[1]).split()[0].split(".")[0] for r in chunk)
        st = collections.Counter()
        for r in chunk:
            for i in stall_cols:
                st[hdr[i][6:]] += int(r[i] or 0)
        ex = sum(int(r[ii]) for r in chunk)
        print(f"[{b0:5d}] {100 * s_ / tot:5.1f}% exec={ex / 1e6:7.2f}M ops={dict(ops.most_common(5))} stalls={dict(st.most_common(4))}")
    break   # launches of one capture are alike: the first is enough
