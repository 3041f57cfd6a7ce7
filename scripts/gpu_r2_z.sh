#!/bin/bash
# round 2, visit Z (1 GPU): K-multi, contiguous batches with incremental jumps against round-robin batches
mkdir -p gpurun_out; L=gpurun_out/r2z.log; rm -f $L
for V in 0 1 0 1; do
CF_MULTI_STRIDED=$V timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_cfg4.json 2> gpurun_out/r2z_cfg4.err
python - gpurun_out/r2z_cfg4.json $V >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("strided", sys.argv[2], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"])
PY
done
cat $L
