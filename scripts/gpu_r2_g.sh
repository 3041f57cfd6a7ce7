#!/bin/bash
# round 2, visit G (2 GPUs): LL exchange protocol -- multi-device / multi-process tests, bench at 2, config 5 at 2
mkdir -p gpurun_out; L=gpurun_out/r2g.log; rm -f $L
timeout 420 python -m pytest tests -m gpu -x -q -k "multidev or peers or sessions" 2>&1 | tail -8 >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err
tail -3 gpurun_out/r2g_bench2.err >> $L
for c in 5 2 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config $c --steps 5 --warmup 3 > gpurun_out/r2g_cfg${c}_2.json 2> gpurun_out/r2g_cfg${c}_2.err
tail -3 gpurun_out/r2g_cfg${c}_2.err >> $L
done
timeout 300 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_cfg5_1.json 2> gpurun_out/r2g_cfg5_1.err
cat $L; for f in gpurun_out/r2g_bench2.json gpurun_out/r2g_cfg5_2.json gpurun_out/r2g_cfg2_2.json gpurun_out/r2g_cfg4_2.json gpurun_out/r2g_cfg5_1.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["config"]["workload"][:40], "n", d["n_gpus"], "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "kms", d["roofline"].get("kernel_ms"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
