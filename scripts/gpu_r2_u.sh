#!/bin/bash
# round 2, visit U (1 GPU): full suite on the final build (new tests: BS fast path large runs, mcSimulAADMulti payoffs), config 2 line
mkdir -p gpurun_out; L=gpurun_out/r2u.log; rm -f $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/r2u_cfg2.json 2> gpurun_out/r2u_cfg2.err
python - gpurun_out/r2u_cfg2.json >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
PY
cat $L
