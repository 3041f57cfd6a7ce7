#!/bin/bash
# what the driver does at round end, in one visit: smoke(), pytest -m gpu, bench (ours + reference arm)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print({k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-200
