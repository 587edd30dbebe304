#!/usr/bin/env python
"""profiles/sass_mix.py — dynamic SASS opcode mix of a kernel from an ncu --set full capture
(`--import-source on`):  python profiles/sass_mix.py <file.ncu-rep> <particles per launch> [top]"""
import collections
import csv
import re
import subprocess
import sys

rep, npart = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
secs, hdr, sec, name = [], None, [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        if sec:
            secs.append((name, hdr, sec))
        sec, hdr, name = [], None, r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr:
        sec.append(r)
if sec:
    secs.append((name, hdr, sec))
name, hdr, sec = secs[0]
ie, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie]) for r in sec)
print(f"{name[:100]}\n{len(sec)} SASS instructions, {tot} warp-instructions executed = {tot * 32 / npart:.1f} per particle")
ops, stalls = collections.Counter(), collections.Counter()
for r in sec:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[ie])
    stalls[op] += int(r[ist])
for op, c in ops.most_common(top):
    print(f"  {op:10s} {c * 32 / npart:7.2f} /particle   stall samples {stalls[op]}")
