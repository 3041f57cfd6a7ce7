"""Run the secondary configs (2: BS x UOC, 5: DLM x Autocall) through the host API a few times: target for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import time
from compfinance_b200.api import CompFinance
import bench_configs as B

which = sys.argv[1] if len(sys.argv) > 1 else "5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
cf = CompFinance(device=0)
B.setup(cf)
for it in range(3):
    t0 = time.perf_counter()
    if which == "5":
        cf.aad_risk_one("dlm5", "auto5", n, sobol=False)
    elif which == "5v":
        cf.value("dlm5", "auto5", n, sobol=False)
    elif which == "2":
        cf.aad_risk_one("bs", "uoc2", n)
    print(which, n, "ms", 1e3 * (time.perf_counter() - t0))
