#!/bin/bash
# round 2, visit T (1 GPU): Black-Scholes reverse kernel, warps per block experiment (build-time CF_BS_REV_WARPS)
mkdir -p gpurun_out; L=gpurun_out/r2t.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_europeans.py -m gpu -x -q 2>&1 | tail -4 >> $L
for i in 1 2; do
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_cfg2.json 2> gpurun_out/r2t_cfg2.err
python - gpurun_out/r2t_cfg2.json >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"])
PY
done
cat $L
