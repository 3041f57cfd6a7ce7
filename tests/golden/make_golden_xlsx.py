#!/usr/bin/env python3
"""Extract the reference's own shipped golden vectors (cached XLL outputs of AutocallPricer.xlsx and
testDLM.xlsx) with their inputs into tests/golden/golden_xlsx.json.  Run in the build
container (needs /root/reference); the JSON is committed."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from xlsx_cells import read_sheet, block  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
G = {"generator": "tests/golden/make_golden_xlsx.py", "source": "cached XLL outputs shipped with asavine/CompFinance"}


def col(cells, a, b):
    return [r[0] for r in block(cells, a, b)]


# ---- AutocallPricer.xlsx, sheet Pricer: xPutDLM (K2), xPutAutocall (K3), xValue (R15), xAADrisk (Q19:R45)
c = read_sheet(os.path.join(REF, "AutocallPricer.xlsx"), 0)
n = int(c["K8"][0])
nd = int(c["K37"][0])
G["autocall_pricer"] = dict(
    spots=col(c, "D9", f"D{8+n}"), atms=col(c, "E9", f"E{8+n}"), skews=col(c, "F9", f"F{8+n}"),
    disc_rate=c["E52"][0], repo_spreads=[(v or 0.0) for v in col(c, "H53", f"H{52+n}")],
    div_dates=col(c, "C40", f"C{39+nd}"),
    divs=[[(v or 0.0) for v in row] for row in block(c, "D40", f"{chr(ord('D')+n-1)}{39+nd}")],
    correl=[[(v or 0.0) for v in row] for row in block(c, "D23", f"{chr(ord('D')+n-1)}{22+n}")],
    lam=c["K34"][0],
    maturity=c["U4"][0], periods=int(c["U5"][0]), ko=c["U6"][0], strike=c["U7"][0], cpn=c["U8"][0], smooth=c["U9"][0],
    n_paths=int(c["U11"][0]), sobol=True,
    payoff_label=c["Q15"][0], price=c["R15"][0],
    risk_labels=col(c, "Q20", "Q45"), risks=[(v or 0.0) for v in col(c, "R20", "R45")], risk_value=c["R19"][0])

# ---- testDLM.xlsx, second sheet: xPutDLM (D16), xPutMultiStats (T3), xValue (T16:U42), 10^6 Sobol paths
c = read_sheet(os.path.join(REF, "testDLM.xlsx"), 1)
G["test_dlm"] = dict(
    spots=block(c, "D8", "F8")[0], atms=block(c, "D9", "F9")[0], skews=block(c, "D10", "F10")[0], disc_rate=c["D5"][0],
    repo_spreads=block(c, "H4", "J4")[0], div_dates=col(c, "H7", "H8"), divs=block(c, "I7", "K8"),
    correl=block(c, "D11", "F13"), lam=c["D6"][0],
    fix_dates=col(c, "X3", "X4"), fwd_dates=col(c, "Y3", "Y4"), n_paths=int(c["U13"][0]), sobol=True,
    labels=col(c, "T16", "T42"), values=col(c, "U16", "U42"))

with open(os.path.join(HERE, "golden_xlsx.json"), "w") as fh:
    json.dump(G, fh, indent=1)
print(json.dumps(G["autocall_pricer"], indent=0)[:1500])
print(G["test_dlm"]["labels"][:3], G["test_dlm"]["values"][:3])
