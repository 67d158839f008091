#!/usr/bin/env python
"""Instruction evidence from the SASS of the shipped library: per kernel, how many of the Blackwell-specific instructions it contains
(static counts from `cuobjdump -sass fl-slam_b200/libgcl_b200.so`), plus the lines themselves for the tensor-core kernel.

    python tools/sass_evidence.py > profiles/r2_sass_evidence.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fl-slam_b200", "libgcl_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "LDGSTS", "REDUX", "CREDUX", "FMNMX3", "FFMA2", "FADD2", "MUFU.EX2", "F2FP.BF16", "DFMA", "DADD", "HMMA", "DMMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, lines = None, collections.OrderedDict(), collections.defaultdict(list)
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(gcl::ScanParams\)|gcl::|void ", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + ".") or (w.endswith(".EX2") and op == w) or (w == "F2FP.BF16" and op.startswith("F2FP.BF16")):
                counts[kern][w] += 1
                if w in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP") and len(lines[(kern, w)]) < 3:
                    lines[(kern, w)].append(re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]+\*/", "", ln)).strip())
print("# SASS evidence (static instruction counts per kernel of `libgcl_b200.so`, sm_100a)\n")
print("Made by `tools/sass_evidence.py` from `cuobjdump -sass`; kernels without any of the listed instructions are left out.\n")
print("| kernel | " + " | ".join(WANT) + " |")
print("|---|" + "---|" * len(WANT))
for k, c in counts.items():
    if sum(c.values()):
        print(f"| `{k}` | " + " | ".join(str(c[w]) if c[w] else "" for w in WANT) + " |")
print("\nUTCHMMA = `tcgen05.mma` (5th-generation tensor core, accumulators in tensor memory), UTCBAR = `tcgen05.commit`, LDTM = `tcgen05.ld`,")
print("UBLKCP = `cp.async.bulk` (bulk copy engine / TMA, mbarrier completion), SYNCS = mbarrier operations, LDGSTS = `cp.async`,")
print("REDUX / CREDUX = warp-wide integer sum / float max in one instruction, FMNMX3 = three-input max, FFMA2 / FADD2 = packed fp32 pairs.\n")
print("## Sample lines\n")
for (k, w), ls in lines.items():
    print(f"* `{k}`, {w}:")
    for l in ls:
        print(f"  * `{l}`")
