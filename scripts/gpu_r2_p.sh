#!/bin/bash
# round 2, visit P (1 GPU): displaced-model kernel v2 with true REDs -- 12 against 8 warps
mkdir -p gpurun_out; L=gpurun_out/r2p.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -5 >> $L
for W in 8 12; do
CF_DLM_WARPS=$W timeout 300 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2p_cfg5_$W.json 2> gpurun_out/r2p_cfg5.err
tail -3 gpurun_out/r2p_cfg5.err >> $L
python - gpurun_out/r2p_cfg5_$W.json >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dlm_kernel -s 1 -c 1 -o gpurun_out/r2p_dlm python scripts/prof_configs.py 5 1048576 > gpurun_out/r2p_ncu_dlm.log 2>&1
tail -3 gpurun_out/r2p_ncu_dlm.log >> $L
cat $L
