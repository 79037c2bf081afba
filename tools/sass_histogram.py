#!/usr/bin/env python
"""SASS opcode histogram of one kernel of libkmerjs_b200.so (static: instructions in the binary, not executed counts).
usage: tools/sass_histogram.py <mangled-name-substring> [lib]"""
import collections
import re
import subprocess
import sys

pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "kmerjs_b200/libkmerjs_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ops, names = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if pat in m.group(1) else None
        if cur:
            names.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops[cur][m.group(2)] += 1
for name in names:
    c = ops[name]
    total = sum(c.values())
    base = collections.Counter()
    for op, n in c.items():
        base[op.split(".")[0]] += n
    print(f"== {name}: {total} SASS instructions")
    print("   by base opcode: " + ", ".join(f"{op} {n}" for op, n in base.most_common(24)))
    for key in ("UTMALDG", "SYNCS", "LDS", "STS", "ATOMS", "ATOMG", "RED", "LDG", "STG", "SHFL", "LOP3", "SHF", "PRMT", "IMAD", "IDP", "BAR", "BSSY", "CALL"):
        full = {op: n for op, n in c.items() if op.split(".")[0] == key}
        if full:
            print(f"   {key}: " + ", ".join(f"{op} {n}" for op, n in sorted(full.items(), key=lambda kv: -kv[1])[:8]))
