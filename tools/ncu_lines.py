#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: instructions executed and stall samples.
usage: tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + __import__("os").environ["NCU_KERNEL"]] if "NCU_KERNEL" in __import__("os").environ else []) + ["--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; res = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": continue
    if hdr and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        # duplicate 'Source' header: first = cuda source
        try:
            inst = int(d.get("Instructions Executed", "0") or 0)
            samp = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        res.append((fname, int(r[0]), r[1].strip(), inst, samp))
tot_i = sum(x[3] for x in res); tot_s = sum(x[4] for x in res)
print(f"total inst {tot_i:.4g} samples {tot_s}")
print("--- by instructions")
for f, ln, src, inst, samp in sorted(res, key=lambda x: -x[3])[:top]:
    print(f"{100*inst/tot_i:5.1f}% inst {100*samp/max(tot_s,1):5.1f}% smp  {f}:{ln}  {src[:100]}")
print("--- by samples")
for f, ln, src, inst, samp in sorted(res, key=lambda x: -x[4])[:top]:
    print(f"{100*inst/tot_i:5.1f}% inst {100*samp/max(tot_s,1):5.1f}% smp  {f}:{ln}  {src[:100]}")
