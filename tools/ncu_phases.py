#!/usr/bin/env python3
"""Instruction / stall-sample share per source-line range of cf_dupire.cuh. usage: tools/ncu_phases.py rep name:lo-hi ..."""
import csv, subprocess, sys
rep = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    nm, r = a.split(":"); f, r = (r.split("@") + [None])[:2] if "@" in r else (r, None)
    lo, hi = f.split("-"); ranges.append((nm, int(lo), int(hi), r))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + __import__("os").environ["NCU_KERNEL"]] if "NCU_KERNEL" in __import__("os").environ else []) + ["--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; res = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        try: res.append((fname, int(r[0]), int(d.get("Instructions Executed", "0") or 0), int(d.get("# Samples", "0") or 0)))
        except ValueError: pass
ti = sum(x[2] for x in res); ts = sum(x[3] for x in res)
byfile = {}
for f, ln, i, s in res:
    byfile.setdefault(f, [0, 0]); byfile[f][0] += i; byfile[f][1] += s
for f, (i, s) in byfile.items(): print(f"file {f:28s} inst {100*i/ti:5.1f}%  smp {100*s/ts:5.1f}%")
for nm, lo, hi, f in ranges:
    i = sum(x[2] for x in res if lo <= x[1] <= hi and x[0] == (f or "cf_dupire.cuh")); s = sum(x[3] for x in res if lo <= x[1] <= hi and x[0] == (f or "cf_dupire.cuh"))
    print(f"{nm:20s} {lo}-{hi}: inst {100*i/ti:5.1f}% ({i/ (32768*156):6.1f}/step-warp)  smp {100*s/ts:5.1f}%")
