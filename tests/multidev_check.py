"""Run by tests/test_gpu_multidev.py in a fresh process with N >= 2 visible GPUs: the same entry points on a one-device
context and on a single-process N-device context (cf_init with a device list: paths sharded over the devices, rank sum
over peer memory inside the reduction kernels) must agree; the multi-device results must repeat bit for bit."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compfinance_b200.api import CompFinance   # noqa: E402

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def workloads(cf):
    out = {}
    spots = np.arange(55, 201, 5.0)
    times = np.arange(1, 37) / 12.0
    vols = 0.15 + 0.10 * np.log(spots[:, None] / 100.0) ** 2 + 0.02 * times[None, :]
    cf.put_dupire(100.0, spots, times, vols, 0.25, "dup")
    cf.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "uoc")
    # north star, a ragged path count (the last shard takes the remainder)
    v, d, vega = cf.dupire_aad_risk("dup", "uoc", [0.7, 0.3], 30, 36, 100_003)
    out["dupire_uoc"] = np.concatenate([[v, d], vega.ravel()])
    out["dupire_uoc_value"] = cf.value("dup", "uoc", 100_003)
    # generic kernel + itemised risk by strike class (config 4 shape, mrg32k3a)
    mats = np.repeat(0.25 * np.arange(1, 13), 60)
    ks = np.tile(70.5 + np.arange(60), 12)
    cf.put_europeans(mats, ks, "euros")
    r = cf.aad_risk_multi("dup", "euros", 1 << 15, sobol=False)
    out["multi_values"], out["multi_risks"] = np.asarray(r[0]), np.asarray(r[1]).ravel()
    nots = np.zeros(720); nots[::7] = 1.0
    a = cf.aad_risk_aggregate("dup", "euros", nots, 1 << 15, sobol=False)
    out["euros_agg"] = np.concatenate([np.asarray(a[0]), [a[1]], np.asarray(a[2])])
    # Black-Scholes barrier (config 2 shape)
    cf.put_black_scholes(100.0, 0.15, False, 0.03, 0.01, "bs")
    cf.put_barrier(100.0, 120.0, 1.0, 1.0 / 52, 0.01, False, "uoc1")
    b = cf.aad_risk_one("bs", "uoc1", 1 << 16)
    out["bs_uoc"] = np.concatenate([np.asarray(b[0]), [b[1]], np.asarray(b[2])])
    # displaced multi-asset autocall (config 5 shape)
    A = 10
    sp = 100.0 + 5.0 * np.arange(A)
    atm = 0.20 + 0.02 * np.arange(A)
    skew = np.array([0.0 if a % 3 == 0 else -0.05 * (a % 3) for a in range(A)])
    repo = 0.001 * np.arange(A)
    corr = np.full((A, A), 0.5); np.fill_diagonal(corr, 1.0)
    cf.put_displaced(sp, atm, skew, 0.02, repo, [0.5, 1.5], np.full((2, A), 0.01), corr, 0.25, "dlm")
    cf.put_autocall(sp, 3.0, 12, 1.0, 0.7, 0.10, 0.01, "auto")
    c = cf.aad_risk_one("dlm", "auto", 1 << 16, sobol=False)
    out["dlm_autocall"] = np.concatenate([np.asarray(c[0]), [c[1]], np.asarray(c[2])])
    # per-path outputs are gathered over the devices in path order
    out["paths"] = np.asarray(cf.simul_paths("dup", "uoc", 5000)).ravel()
    return out


single = workloads(CompFinance(device=0))
cfm = CompFinance(devices=list(range(n_dev)))
multi = workloads(cfm)
again = workloads(cfm)
report = {}
for k in single:
    a, b, c = single[k], multi[k], again[k]
    scale = np.maximum(np.abs(a), 1e-6 * max(1.0, float(np.max(np.abs(a)))))
    report[k] = {"max_rel": float(np.max(np.abs(a - b) / scale)), "bitwise_repeat": bool(np.array_equal(b, c)), "n": int(a.size)}
print("MULTIDEV " + json.dumps(report))
