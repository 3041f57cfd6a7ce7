"""Run the config-3 (Dupire x UOC, 156 steps, 30x36 surface) AAD kernel a few times: target for ncu."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from compfinance_b200 import capi
from oracle import restate as R

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "aad"
eng = capi.Engine(device=0)
spots = np.arange(55, 201, 5.0); times = np.arange(1, 37) / 12.0
vols = 0.15 + 0.10 * np.log(spots[:, None] / 100) ** 2 + 0.02 * times[None, :]
ptl = R.uoc_timeline(3.0, 1.0 / 52)
tab = R.DupireTables(100, spots, times, vols, 0.25, ptl)
mdl = eng.dupire_model(100.0, tab.log_spots, tab.interp_vols, tab.common, len(ptl), time_map=tab.time_map())
prd = eng.uoc(120.0, 150.0, float(np.exp(np.log(100.0)) * 0.01), len(ptl))
if os.environ.get("CF_PROF_BARRIER"):
    prd = eng.uoc(120.0, float(os.environ["CF_PROF_BARRIER"]), float(np.exp(np.log(100.0)) * 0.01), len(ptl))
rg = eng.rng("sobol")
plan = C.c_void_p()
eng._chk(eng.lib.cf_plan_create(C.byref(mdl), C.byref(prd), C.byref(rg), C.byref(plan)))
dout = torch.zeros(eng.lib.cf_plan_out_size(plan, 1), dtype=torch.float64, device="cuda")
wv = (C.c_double * 2)(1.0, 0.0)
stream = torch.cuda.Stream() if os.environ.get("CF_PROF_STREAM") else None
sp = C.c_void_p(stream.cuda_stream) if stream else None
first = int(os.environ.get("CF_PROF_FIRST", "0"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
step_ms = []
for it in range(iters):
    flush.fill_(it & 1)                     # L2 flush between iterations (untimed)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if mode == "aad":
        eng._chk(eng.lib.cf_plan_launch_aad(plan, wv, first, N, dout.data_ptr(), sp))
    else:
        eng._chk(eng.lib.cf_plan_launch_value(plan, first, N, dout.data_ptr(), sp))
    e1.record(stream)
    torch.cuda.synchronize()
    step_ms.append(e0.elapsed_time(e1))
torch.cuda.synchronize()
step_ms = sorted(step_ms[2:]) if len(step_ms) > 4 else step_ms
print("step ms (incl. reduction) median %.4f min %.4f" % (step_ms[len(step_ms) // 2], step_ms[0]), "sum %.17g" % float(dout[:4].sum().item()))
ms = C.c_double(); nl = C.c_int()
eng._chk(eng.lib.cf_plan_kernel_ms(plan, C.byref(ms), C.byref(nl)))
print(mode, "kernel avg ms", ms.value, "paths/s %.4g" % (N / ms.value * 1e3), "fp64 peak TF", eng.fp64_peak_tflops())
if os.environ.get("CF_DEBUG_TIMES"):
    buf = np.zeros((3, 1024, 8), dtype=np.uint64)
    eng._chk(eng.lib.cf_plan_debug_times(plan, buf.ctypes.data_as(C.c_void_p)))
    nb = 148
    f, r, lv = buf[0, :nb].astype(np.int64), buf[1, :nb].astype(np.int64), buf[2, :nb, 0].astype(np.int64)
    t0 = f[:, 0].min()
    def stat(name, x):
        print("  %-34s min %8.1f  mean %8.1f  max %8.1f us" % (name, x.min() / 1e3, x.mean() / 1e3, x.max() / 1e3))
    stat("fwd entry (since first entry)", f[:, 0] - t0)
    stat("fwd tables staged", f[:, 1] - f[:, 0])
    stat("fwd units done", f[:, 2] - f[:, 1])
    stat("fwd end (since first entry)", f[:, 3] - t0)
    if r[:, 0].max() > 0:
        stat("rev entry (since fwd first entry)", r[:, 0] - t0)
        stat("rev tables staged", r[:, 1] - r[:, 0])
        stat("rev dependency wait", r[:, 2] - r[:, 1])
        stat("rev live compaction", r[:, 3] - r[:, 2])
        stat("rev sweep", r[:, 4] - r[:, 3])
        stat("rev block sum + barrier", r[:, 5] - r[:, 4])
        stat("rev combine tables", r[:, 6] - r[:, 5])
        stat("rev end (since fwd first entry)", r[:, 6] - t0)
        print("  live paths per block: min %d mean %.1f max %d" % (lv.min(), lv.mean(), lv.max()))
