#!/bin/bash
# round 2, visit L (1 GPU): Black-Scholes fast path -- parity + timing of configs 1, 2
mkdir -p gpurun_out; L=gpurun_out/r2l.log; rm -f $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L
for c in 2 1; do
timeout 300 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2l_cfg${c}.json 2> gpurun_out/r2l_cfg${c}.err
tail -2 gpurun_out/r2l_cfg${c}.err >> $L
python - gpurun_out/r2l_cfg${c}.json >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
for N in 131072; do
  for V in "CF_DUPIRE_REV=classic" "CF_DUPIRE_REV=span"; do
    echo "== N=$N $V" >> $L
    env $V CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|rev live|rev end" >> $L
  done
done
cat $L
