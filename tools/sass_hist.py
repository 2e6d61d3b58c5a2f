#!/usr/bin/env python
"""Opcode histogram per kernel from `cuobjdump -sass` (developer tool): FP64 vs everything else."""
import collections
import re
import subprocess
import sys

obj = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn = None
hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
for fn, h in hist.items():
    if pat and pat not in fn:
        continue
    tot = sum(h.values())
    fp64 = sum(v for k, v in h.items() if k in ("DFMA", "DADD", "DMUL"))
    print(f"{fn}: total {tot}  fp64 {fp64} ({100.0 * fp64 / tot:.1f}%)")
    print("   " + "  ".join(f"{k}:{v}" for k, v in h.most_common(18)))
