#!/usr/bin/env python3
"""Stall reasons of an ncu report, overall and per CUDA source line (warp-state samples of the source page).
usage: tools/ncu_stalls.py report.ncu-rep [top]"""
import collections, csv, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if "NCU_KERNEL" in os.environ:
    cmd += ["--kernel-name", "regex:" + os.environ["NCU_KERNEL"]]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
hdr = None; cur = None
tot = collections.Counter(); per = collections.defaultdict(collections.Counter); inst = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "" and r[2] == "-":
        key = (cur, int(r[0]), r[1].strip()[:90])
        try: inst[key] += int(r[7])
        except ValueError: pass
        for i, name in enumerate(hdr):
            if name.startswith("stall_") and "Not Issued" not in name:
                try: v = int(r[i])
                except ValueError: continue
                tot[name] += v; per[key][name] += v
n = sum(tot.values()) or 1
print("stall samples by reason")
for k, v in tot.most_common(10): print(f"  {k[6:]:22s} {100 * v / n:5.1f}%")
print("lines by stall samples: share, instructions share, top two reasons")
ni = sum(inst.values()) or 1
for key, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    t = sum(c.values()); m = c.most_common(2)
    why = ", ".join(f"{a[6:]} {100 * b / t:.0f}%" for a, b in m)
    print(f"  {100 * t / n:5.2f}%  inst {100 * inst[key] / ni:5.2f}%  {key[0]}:{key[1]}  [{why}]  {key[2]}")
