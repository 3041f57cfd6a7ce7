#!/usr/bin/env python3
"""Shared-memory wavefronts per CUDA source line of an ncu report (source page): total, ideal, excessive (bank conflicts).
usage: NCU_KERNEL=regex tools/ncu_smem.py report.ncu-rep [top]"""
import collections, csv, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if "NCU_KERNEL" in os.environ:
    cmd += ["--kernel-name", "regex:" + os.environ["NCU_KERNEL"]]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
hdr = None; cur = None; per = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {n: i for i, n in enumerate(hdr)}; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "" and r[2] == "-":
        key = (cur, int(r[0]), r[1].strip()[:90])
        def g(name):
            try: return int(r[ix[name]])
            except (KeyError, ValueError): return 0
        v = per[key]
        v[0] += g("L1 Wavefronts Shared"); v[1] += g("L1 Wavefronts Shared Ideal"); v[2] += g("L1 Wavefronts Shared Excessive"); v[3] += g("Instructions Executed")
tot = sum(v[0] for v in per.values()) or 1
print("shared-memory wavefronts: total %d, ideal %d, excessive %d" % (tot, sum(v[1] for v in per.values()), sum(v[2] for v in per.values())))
for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {100 * v[0] / tot:5.2f}%  wf {v[0]:9d} ideal {v[1]:9d} excess {v[2]:9d} inst {v[3]:9d}  {key[0]}:{key[1]}  {key[2]}")
