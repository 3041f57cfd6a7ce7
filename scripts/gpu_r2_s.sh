#!/bin/bash
# round 2, visit S (1 GPU): host chain rule without the tape sweep when the device adjoints land on parameters
mkdir -p gpurun_out; L=gpurun_out/r2s.log; rm -f $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $L
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - gpurun_out/r2s_bench.json >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("bench value %.4g ms %.4f e2e %.4g first_call %.4g frac %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["first_call"]["value"], d["roofline"]["frac"]))
PY
CF_TIMING=1 timeout 300 python scripts/e2e_breakdown.py 2>&1 | tail -12 >> $L
for c in 4; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2s_cfg${c}.json 2> gpurun_out/r2s_cfg${c}.err
python - gpurun_out/r2s_cfg${c}.json >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
PY
done
cat $L
