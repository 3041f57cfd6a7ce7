"""Where the end-to-end call spends its time: resident plan launch vs C-ABI one-shot vs host API (dupireAADRisk)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from compfinance_b200 import capi
from compfinance_b200.api import CompFinance

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cf = CompFinance(device=0)
eng = capi.Engine()
spots = np.arange(55, 201, 5.0); times = np.arange(1, 37) / 12.0
vols = 0.15 + 0.10 * np.log(spots[:, None] / 100) ** 2 + 0.02 * times[None, :]
cf.put_dupire(100.0, spots, times, vols, 0.25, "m")
cf.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "p")
d = cf.describe("m", "p", aad=True)
mdl = eng.dupire_model(100.0, d["tab_b"], d["tab_a"], d["is_event"], d["n_events"], time_map=d["time_map"])
prd = eng.uoc(d["strike"], d["barrier"], d["smooth"], d["n_events"])
rng = eng.rng("sobol")
plan = C.c_void_p()
eng._chk(eng.lib.cf_plan_create(C.byref(mdl), C.byref(prd), C.byref(rng), C.byref(plan)))
dout = torch.zeros(eng.lib.cf_plan_out_size(plan, 1), dtype=torch.float64, device="cuda")
wv = (C.c_double * 2)(1.0, 0.0)

def timeit(f, n=10):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def plan_launch():
    eng._chk(eng.lib.cf_plan_launch_aad(plan, wv, 0, N, dout.data_ptr(), None)); torch.cuda.synchronize()
print("plan launch + sync        ms", timeit(plan_launch))
print("C ABI cf_run_aad          ms", timeit(lambda: eng.run_aad(mdl, prd, rng, 0, N, [1.0, 0.0])))
print("host API dupireAADRisk    ms", timeit(lambda: cf.dupire_aad_risk("m", "p", [1.0, 0.0], 30, 36, N)))
print("host API value            ms", timeit(lambda: cf.value("m", "p", N)))
