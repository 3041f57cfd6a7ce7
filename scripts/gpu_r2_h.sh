#!/bin/bash
# round 2, visit H (1 GPU): span v3 + lean generic Gaussians / DLM -- parity, timing at 2^17 / 2^18 / 2^20, configs 2 and 5
mkdir -p gpurun_out; L=gpurun_out/r2h.log; rm -f $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 >> $L
for N in 131072 262144; do
  for V in "CF_DUPIRE_REV=classic" "CF_DUPIRE_REV=span"; do
    echo "== N=$N $V" >> $L
    env $V CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|rev end|fwd end" >> $L
  done
done
echo "== N=1048576 span" >> $L
CF_DUPIRE_REV=span timeout 300 python scripts/prof_config3.py 1048576 20 aad 2>&1 | tail -2 >> $L
for c in 2 5; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_cfg${c}.json 2> gpurun_out/r2h_cfg${c}.err
tail -2 gpurun_out/r2h_cfg${c}.err >> $L
python - gpurun_out/r2h_cfg${c}.json >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
CF_DUPIRE_REV=span timeout 900 ncu --set full --clock-control none --import-source on -k regex:dupire_reverse_span -s 4 -c 1 -o gpurun_out/r2h_revs_small python scripts/prof_config3.py 131072 6 aad > gpurun_out/r2h_ncu.log 2>&1
cat $L
