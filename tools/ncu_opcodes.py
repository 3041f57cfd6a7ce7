#!/usr/bin/env python3
"""Executed-instruction histogram by SASS opcode. usage: tools/ncu_opcodes.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections, re
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"] + (["--kernel-name", "regex:" + __import__("os").environ["NCU_KERNEL"]] if "NCU_KERNEL" in __import__("os").environ else []) , capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
h = collections.Counter()
for r in rows[2:]:
    if len(r) <= iex: continue
    src = r[isrc].strip()
    mo = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not mo: continue
    op = mo.group(2).split(".")[0]
    try: h[op] += int(r[iex])
    except ValueError: pass
tot = sum(h.values())
print("total", tot)
for op, n in h.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print(f"{op:12s} {n:14d} {100*n/tot:5.1f}%")
