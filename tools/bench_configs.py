#!/usr/bin/env python3
"""Secondary timings: the parity configs of BASELINE.json (1, 2, 4, 5) at their full sizes, end to end through the
reference-facing host API (host buffers in and out, every call), beside the reference's own CPU path on a bounded
sample of the same workload.  Not bench lines (bench.py measures config 3): a record for profiles/.

    python tools/bench_configs.py [--reps 5] [--no-ref] > profiles/rN_configs.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def surface():
    spots = np.arange(55, 201, 5.0)
    times = np.arange(1, 37) / 12.0
    return spots, times, 0.15 + 0.10 * np.log(spots[:, None] / 100.0) ** 2 + 0.02 * times[None, :]


def put_bs(api):
    (api.put_black_scholes if hasattr(api, "put_black_scholes") else api.put_bs)(100.0, 0.15, False, 0.0, 0.0, "bs")


def setup(api):
    spots, times, vols = surface()
    put_bs(api)
    api.put_european(100.0, 1.0, 1.0, "eur")
    api.put_barrier(100.0, 120.0, 1.0, 1.0 / 52, 0.01, False, "uoc2")
    api.put_dupire(100.0, spots, times, vols, 0.25, "dup")
    api.put_europeans(np.repeat(0.25 * np.arange(1, 13), 60), np.tile(70.5 + np.arange(60), 12), "eurs4")
    a = np.arange(10)
    s5 = 100.0 + 5 * a
    api.put_displaced(s5, 0.20 + 0.02 * a, np.where(a % 3 == 0, 0.0, -0.05 * (a % 3)), 0.02, 0.001 * a, [0.5, 1.5],
                      np.full((2, 10), 0.01), np.full((10, 10), 0.5) + 0.5 * np.eye(10), 0.25, "dlm5")
    api.put_autocall(s5, 3.0, 12, 1.0, 0.7, 0.10, 0.01, "auto5")


CASES = [
    # name, call(api, n) -> result, full size, CPU sample size
    ("config1_bs_european_sobol_2^16_aad", lambda a, n: a.aad_risk_one("bs", "eur", n), 1 << 16, 1 << 16),
    ("config2_bs_barrier_sobol_2^20x52_aad", lambda a, n: a.aad_risk_one("bs", "uoc2", n), 1 << 20, 1 << 18),
    ("config4_dupire_europeans_720_mrg_2^22_value", lambda a, n: a.value("dup", "eurs4", n, sobol=False), 1 << 22, 1 << 18),
    ("config4_dupire_europeans_720_mrg_2^22_aad_aggregate",
     lambda a, n: a.aad_risk_aggregate("dup", "eurs4", 0.5 + np.cos(np.arange(720)), n, sobol=False), 1 << 22, 1 << 17),
    ("config4_dupire_europeans_720_mrg_2^22_aad_multi_1081x720", lambda a, n: a.aad_risk_multi("dup", "eurs4", n, sobol=False), 1 << 22, 1 << 12),
    ("config5_dlm10_autocall_mrg_2^22_aad", lambda a, n: a.aad_risk_one("dlm5", "auto5", n, sobol=False), 1 << 22, 1 << 17),
]


def best_of(fn, reps):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    from compfinance_b200.api import CompFinance
    cf = CompFinance(device=0)
    setup(cf)
    ref = None
    if not args.no_ref:
        try:
            from oracle import refapi
            ref = refapi.get()
            threads = ref.start_pool(-1) + 1
            setup(ref)
        except (OSError, FileNotFoundError):
            ref = None
    out = []
    for name, call, n, n_cpu in CASES:
        call(cf, n)                                                   # warm-up (scratch buffers, module load)
        secs = best_of(lambda: call(cf, n), args.reps)
        row = {"config": name, "paths": n, "ms_per_call": 1e3 * secs, "paths_per_sec_e2e": n / secs,
               "api": "libcf_host.so entry point, host buffers, best of %d" % args.reps}
        if ref is not None:
            csecs = best_of(lambda: call(ref, n_cpu), 1)
            row["cpu_reference"] = {"paths_per_sec": n_cpu / csecs, "cores": threads, "sample_paths": n_cpu}
            row["speedup_vs_cpu_reference"] = row["paths_per_sec_e2e"] / (n_cpu / csecs)
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
