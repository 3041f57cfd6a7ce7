#!/usr/bin/env python3
"""Generate tests/golden/golden_r1.json from the reference itself (oracle/_ref/libcfref.so, built by
oracle/build_ref.py from /root/reference).  Run in the build container; the JSON is committed so that
GPU parity tests do not need /root/reference."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import build_ref, refapi  # noqa: E402

build_ref.build("/root/reference", verbose=False)
r = refapi.get()
r.start_pool(-1)
G = {"generator": "tests/golden/make_golden.py", "reference": "asavine/CompFinance compiled with g++ -O3 -march=x86-64-v3"}

# RNG streams
G["sobol_uniforms_dim4_first1000_n4"] = r.rng_draw(True, 4, 1000, 4, False).tolist()
G["sobol_gauss_dim156_first123456_n2"] = r.rng_draw(True, 156, 123456, 2, True).tolist()
G["mrg_uniforms_dim5_first6400_n4"] = r.rng_draw(False, 5, 6400, 4, False).tolist()
G["mrg_gauss_dim12_first64_n3"] = r.rng_draw(False, 12, 64, 3, True).tolist()

# config 1: BS European, Sobol 2^16
r.put_bs(100, 0.15, False, 0.0, 0.0, "bs1"); r.put_european(100, 1.0, 1.0, "eur1")
pv, rv, risks = r.aad_risk_one("bs1", "eur1", 1 << 16)
G["config1"] = dict(n=1 << 16, value=float(r.value("bs1", "eur1", 1 << 16)[0]), risks=risks.tolist())

# config 2: BS UOC 52 weekly steps (r = 3 %, d = 1 %), Sobol and mrg32k3a
r.put_bs(100, 0.15, False, 0.03, 0.01, "bs2"); r.put_barrier(100, 120, 1.0, 1.0 / 52, 0.01, False, "uoc2")
for name, sob, n in [("config2_sobol", True, 1 << 17), ("config2_mrg", False, 1 << 16)]:
    pv, rv, risks = r.aad_risk_one("bs2", "uoc2", n, sobol=sob)
    G[name] = dict(n=n, values=pv.tolist(), risk_value=rv, risks=risks.tolist())
pv, rv, risks = r.aad_risk_one("bs1", "uoc2", 1 << 20)
G["config2_full"] = dict(n=1 << 20, values=pv.tolist(), risks=risks.tolist())     # BASELINE config 2 (r = d = 0)

# config 3: Dupire UOC 156 steps
spots = np.arange(55, 201, 5.0); times = np.arange(1, 37) / 12.0
vols = 0.15 + 0.10 * np.log(spots[:, None] / 100) ** 2 + 0.02 * times[None, :]
r.put_dupire(100, spots, times, vols, 0.25, "dup3"); r.put_barrier(120, 150, 3.0, 1.0 / 52, 0.01, False, "uoc3")
for name, sob, n, w in [("config3_sobol_16k", True, 1 << 14, [1.0, 0.0]), ("config3_mrg_16k", False, 1 << 14, [0.7, 0.3])]:
    val, delta, vega = r.dupire_aad_risk("dup3", "uoc3", w, 30, 36, n, sobol=sob)
    G[name] = dict(n=n, notionals=w, values=r.value("dup3", "uoc3", n, sobol=sob).tolist(), value=val, delta=delta,
                   vega=vega.tolist(), first_paths=r.simul_paths("dup3", "uoc3", 64, sobol=sob).tolist())
val, delta, vega = r.dupire_aad_risk("dup3", "uoc3", [1.0, 0.0], 30, 36, 1 << 20)
G["config3_full"] = dict(n=1 << 20, values=r.value("dup3", "uoc3", 1 << 20).tolist(), value=val, delta=delta, vega=vega.tolist())

# Dupire European with inserted steps, put barrier
r.put_european(110, 1.0, 1.0, "eur110")
val, delta, vega = r.dupire_aad_risk("dup3", "eur110", [1.0], 30, 36, 1 << 14)
G["dupire_european_16k"] = dict(n=1 << 14, value=val, delta=delta, vega=vega.tolist())
r.put_barrier(90, 130, 2.0, 1.0 / 12, 0.02, True, "uop")
val, delta, vega = r.dupire_aad_risk("dup3", "uop", [0.5, 0.5], 30, 36, 1 << 13, sobol=False)
G["dupire_uop_mrg_8k"] = dict(n=1 << 13, value=val, delta=delta, vega=vega.tolist())

with open(os.path.join(HERE, "golden_r1.json"), "w") as fh:
    json.dump(G, fh)
print("wrote golden_r1.json", os.path.getsize(os.path.join(HERE, "golden_r1.json")), "bytes")
print("config3_full", G["config3_full"]["value"], G["config3_full"]["delta"], np.sum(G["config3_full"]["vega"]))
