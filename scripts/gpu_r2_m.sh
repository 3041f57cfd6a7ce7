#!/bin/bash
# round 2, visit M (1 GPU): Black-Scholes fast path in one launch -- parity, timing, launch list of config 2
mkdir -p gpurun_out; L=gpurun_out/r2m.log; rm -f $L
timeout 900 python -m pytest tests -m gpu -x -q -k "europeans or bs or black or odd or generic or sessions or api" 2>&1 | tail -8 >> $L
for c in 2 1; do
timeout 300 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2m_cfg${c}.json 2> gpurun_out/r2m_cfg${c}.err
tail -2 gpurun_out/r2m_cfg${c}.err >> $L
python - gpurun_out/r2m_cfg${c}.json >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2m_cfg2_launches.csv python bench.py --config 2 --steps 2 --warmup 1 > /dev/null 2>&1
python - >> $L <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2m_cfg2_launches.csv")) if len(r)>5]
for r in rows[-14:]:
    print(r[4][:70], r[-1], r[-2])
PY
cat $L
