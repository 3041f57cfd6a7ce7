#!/bin/bash
# round 2, visit V (1 GPU): closing event after the reduction (PDL chain intact) -- shard timing, bench, fast-path tests
mkdir -p gpurun_out; L=gpurun_out/r2v.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sessions.py -m gpu -x -q 2>&1 | tail -3 >> $L
for N in 131072 1048576; do
  echo "== N=$N" >> $L
  CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 30 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|rev end|fwd end" >> $L
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - gpurun_out/r2v_bench.json >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("bench value %.4g ms %.4f e2e %.4g first_call %.4g frac %.3f kernel_ms %.4f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["first_call"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]))
PY
cat $L
