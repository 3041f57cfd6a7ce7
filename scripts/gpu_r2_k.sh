#!/bin/bash
# round 2, visit K (1 GPU): BS reverse rework + span scan -- parity, timings, bench, K-multi ncu summary, launch list
mkdir -p gpurun_out; L=gpurun_out/r2k.log; rm -f $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $L
for N in 131072 196608; do
  for V in "CF_DUPIRE_REV=classic" "CF_DUPIRE_REV=span"; do
    echo "== N=$N $V" >> $L
    env $V CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|rev live|rev end" >> $L
  done
done
for c in 2 4 5 1; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2k_cfg${c}.json 2> gpurun_out/r2k_cfg${c}.err
tail -2 gpurun_out/r2k_cfg${c}.err >> $L
python - gpurun_out/r2k_cfg${c}.json >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err
timeout 600 ncu --set full --clock-control none -k regex:dupire_europeans_multi -s 0 -c 1 -o gpurun_out/r2k_multi python -c "
import sys; sys.path.insert(0,'.')
import bench
from compfinance_b200.api import CompFinance
cf=CompFinance(device=0); m=bench._put_config(cf,4)
for _ in range(2): cf.aad_risk_multi(m,'bench_prd',1<<20,sobol=False)
" > gpurun_out/r2k_ncu_multi.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cat $L; cat gpurun_out/r2k_bench.json
