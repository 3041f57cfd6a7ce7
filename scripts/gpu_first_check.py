"""First GPU check: RNG parity and low-level engine vs the numpy restatement / compiled reference."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from compfinance_b200 import capi
from oracle import restate as R

eng = capi.Engine(device=0)
# direction numbers
d = R.sobol_direction_numbers()
print("dirnum", eng.direction_number(20, 155), d[20, 155])
# sobol states
for (dim, first, n) in [(7, 0, 1000), (156, 123456, 700), (3, 255, 513), (1101, 1 << 20, 300)]:
    a = eng.sobol_states(dim, first, n); b = R.sobol_states(dim, first, n)
    print("sobol states", dim, first, n, "equal:", bool((a == b).all()))
rm = eng.rng("mrg")
for (dim, first, n) in [(3, 0, 10), (12, 6400, 100), (120, 100001, 50)]:
    a = eng.mrg_numerators(rm, dim, first, n); b = R.mrg32k3a_numerators(12345, 12346, dim, first, n)
    print("mrg numerators", dim, first, n, "equal:", bool((a == b).all()))
p = np.concatenate([np.linspace(1e-9, 1 - 1e-9, 100001), [0.5, 0.50000000000000078, 0.08, 0.92]])
a = eng.inv_normal(p); b = R.inv_normal_cdf(p)
print("invnormal max abs diff", np.abs(a - b).max())
g = eng.rng_draw(eng.rng("sobol"), 156, 1000, 64, True); gb = R.gaussians(("sobol",), 156, 1000, 64)
print("sobol gauss maxdiff", np.abs(g - gb).max())
g = eng.rng_draw(rm, 12, 6401, 64, True); gb = R.gaussians(("mrg32k3a", 12345, 12346), 12, 6401, 64)
print("mrg gauss maxdiff", np.abs(g - gb).max())

# Dupire x UOC
spots = np.arange(55, 201, 5.0); times = np.arange(1, 37) / 12.0
vols = 0.15 + 0.10 * np.log(spots[:, None] / 100) ** 2 + 0.02 * times[None, :]
ptl = R.uoc_timeline(3.0, 1.0 / 52)
tab = R.DupireTables(100, spots, times, vols, 0.25, ptl)
mdl = eng.dupire_model(100.0, tab.log_spots, tab.interp_vols, tab.common, len(ptl))
mdlT = eng.dupire_model(100.0, tab.log_spots, tab.interp_vols, tab.common, len(ptl), time_map=tab.time_map())
smooth_abs = float(np.exp(np.log(100.0)) * 0.01)
prd = eng.uoc(120.0, 150.0, smooth_abs, len(ptl))
w = [0.7, 0.3]
for rngname, rt in [("sobol", ("sobol",)), ("mrg", ("mrg32k3a", 12345, 12346))]:
    for (first, N) in [(0, 1 << 14), (12345 * 2, 5000)]:
        o = R.dupire_uoc_run(tab, dict(strike=120, barrier=150, smooth=0.01), rt, first, N, w)
        t = time.time()
        v, pp = eng.run_value(mdl, prd, eng.rng(rngname), first, N, per_path=True)
        r = eng.run_aad(mdl, prd, eng.rng(rngname), first, N, w, per_path=True)
        dt = time.time() - t
        print(f"dupire-uoc {rngname} first={first} N={N}: value per-path maxdiff {np.abs(pp - o['payoffs']).max():.3e}",
              f"sum rel {np.abs(v / o['payoffs'].sum(0) - 1).max():.3e}",
              f"aad per-path {np.abs(r['payoffs'] - o['payoffs']).max():.3e} agg {np.abs(r['agg'] - o['agg']).max():.3e}",
              f"spot_adj rel {abs(r['table_adj'][0] / o['spot_adj'] - 1):.3e}",
              f"ybar max abs {np.abs(r['table_adj'][1:].reshape(tab.n_steps, -1) - o['ybar']).max():.3e} (scale {np.abs(o['ybar']).max():.3e})", f"t={dt:.2f}s")
        rT = eng.run_aad(mdlT, prd, eng.rng(rngname), first, N, w, per_path=True)
        so, vo = tab.param_risks(o['spot_adj'], o['ybar'], N)
        vT = rT['table_adj'][1:].reshape(len(spots), len(times)) / N
        big = np.abs(vo) > 1e-6
        print(f"   fast kernel: per-path {np.abs(rT['payoffs'] - o['payoffs']).max():.3e} agg {np.abs(rT['agg'] - o['agg']).max():.3e}",
              f"delta rel {abs(rT['table_adj'][0] / N / so - 1):.3e} vega max abs {np.abs(vT - vo).max():.3e} max rel(>1e-6) {np.abs(vT[big] / vo[big] - 1).max():.3e}")
# determinism
r1 = eng.run_aad(mdlT, prd, eng.rng("sobol"), 0, 1 << 15, w); r2 = eng.run_aad(mdlT, prd, eng.rng("sobol"), 0, 1 << 15, w)
print("bitwise deterministic:", bool((r1["table_adj"] == r2["table_adj"]).all() and r1["agg_sum"] == r2["agg_sum"]))

# BS x European / UOC
tb = R.BSTables(100, 0.15, 0.03, 0.01, [1.0], [1.25], [1.25], [True])
m1 = eng.bs_model(100.0, tb.drifts, tb.stds, tb.is_event, tb.numeraires, tb.fwd_factors, tb.discounts)
N = 1 << 14
o = R.bs_run(tb, "european", dict(strike=100), ("sobol",), 0, N, [1.0])
r = eng.run_aad(m1, eng.european(100.0), eng.rng("sobol"), 0, N, [1.0], per_path=True)
print("bs-eur per-path", np.abs(r["payoffs"] - o["payoffs"]).max(), "adj rel", np.abs(r["table_adj"] / o["table_adj"] - 1).max())
ptl2 = R.uoc_timeline(1.0, 1.0 / 52)
tb2 = R.BSTables(100, 0.15, 0.03, 0.01, ptl2, ptl2, [None] * len(ptl2), [False] * (len(ptl2) - 1) + [True])
m2 = eng.bs_model(100.0, tb2.drifts, tb2.stds, tb2.is_event, tb2.numeraires, tb2.fwd_factors, tb2.discounts)
for rngname, rt in [("sobol", ("sobol",)), ("mrg", ("mrg32k3a", 12345, 12346))]:
    o = R.bs_run(tb2, "uoc", dict(strike=100, barrier=120, smooth=0.01), rt, 0, N, [1.0, 0.0])
    r = eng.run_aad(m2, eng.uoc(100.0, 120.0, 1.0, len(ptl2)), eng.rng(rngname), 0, N, [1.0, 0.0], per_path=True)
    ra, oa = r["table_adj"], o["table_adj"]
    nz = np.abs(oa) > 1e-12
    print("bs-uoc", rngname, "per-path", np.abs(r["payoffs"] - o["payoffs"]).max(), "adj rel(nz)", np.abs(ra[nz] / oa[nz] - 1).max(), "abs(z)", np.abs(ra[~nz] - oa[~nz]).max() if (~nz).any() else 0,
          "risks", tb2.param_risks(ra, N), tb2.param_risks(oa, N))

# quick timing config 3
import ctypes as C
N = 1 << 20
t = time.time(); r = eng.run_aad(mdlT, prd, eng.rng("sobol"), 0, N, [1.0, 0.0]); print("config3 AAD one-shot wall", time.time() - t)
t = time.time(); r = eng.run_aad(mdlT, prd, eng.rng("sobol"), 0, N, [1.0, 0.0]); print("config3 AAD one-shot wall (2nd)", time.time() - t)
sr, vr = r["table_adj"][0] / N, r["table_adj"][1:].reshape(len(spots), len(times)) / N
print("value %.17g delta %.17g sumvega %.17g vega[13][11] %.17g" % (r["agg_sum"] / N, sr, vr.sum(), vr[13][11]))
print("ref:  0.96926107424976005 0.020804057371458962 -5.4262407757753115 -0.024065608611340886")
plan = C.c_void_p()
rg = eng.rng("sobol")
eng._chk(eng.lib.cf_plan_create(C.byref(mdlT), C.byref(prd), C.byref(rg), C.byref(plan)))
import torch
nout = eng.lib.cf_plan_out_size(plan, 1)
dout = torch.zeros(nout, dtype=torch.float64, device="cuda")
wv = (C.c_double * 2)(1.0, 0.0)
for it in range(5):
    eng._chk(eng.lib.cf_plan_launch_aad(plan, wv, 0, N, dout.data_ptr(), None))
torch.cuda.synchronize()
ms = C.c_double(); nl = C.c_int()
eng._chk(eng.lib.cf_plan_kernel_ms(plan, C.byref(ms), C.byref(nl)))
print("AAD kernel avg ms", ms.value, "launches", nl.value, "paths/s", N / ms.value * 1e3)
dv = torch.zeros(2, dtype=torch.float64, device="cuda")
for it in range(5):
    eng._chk(eng.lib.cf_plan_launch_value(plan, 0, N, dv.data_ptr(), None))
torch.cuda.synchronize()
eng._chk(eng.lib.cf_plan_kernel_ms(plan, C.byref(ms), C.byref(nl)))
print("value kernel avg ms", ms.value, "launches", nl.value, "paths/s", N / ms.value * 1e3, dv.cpu().numpy() / N)
