#!/bin/bash
# round 2, visit R (1 GPU): call-ladder payoff sums (one pass over shared forwards instead of a shuffle tree per strike)
mkdir -p gpurun_out; L=gpurun_out/r2r.log; rm -f $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $L
for c in 4; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2r_cfg${c}.json 2> gpurun_out/r2r_cfg${c}.err
tail -2 gpurun_out/r2r_cfg${c}.err >> $L
python - gpurun_out/r2r_cfg${c}.json >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 600 python tools/bench_configs.py > gpurun_out/r2r_configs.json 2> gpurun_out/r2r_configs.err; tail -12 gpurun_out/r2r_configs.err >> $L
cat $L
