#!/bin/bash
# round 2, visit X (1 GPU): K-multi with one call site for the sweeps (code 190 KB -> 52 KB) -- parity, config 4 timing
mkdir -p gpurun_out; L=gpurun_out/r2x.log; rm -f $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $L
for c in 4; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2x_cfg${c}.json 2> gpurun_out/r2x_cfg${c}.err
python - gpurun_out/r2x_cfg${c}.json >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
PY
done
timeout 600 python tools/bench_configs.py > gpurun_out/r2x_configs.json 2> gpurun_out/r2x_configs.err
python - >> $L <<'PY'
import json
for l in open('gpurun_out/r2x_configs.json'):
    l=l.strip()
    if l:
        d=json.loads(l); print(d['config'], 'ms %.4g'%d['ms_per_call'], 'x cpu %.4g'%d['speedup_vs_cpu_reference'])
PY
cat $L
