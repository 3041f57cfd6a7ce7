#!/bin/bash
# round 2, final 1-GPU visit: full GPU suite, smoke, default bench (ours + reference arm), config lines, launch list,
# (the ncu --set full captures are in gpu_r2_final_ncu.sh: the reports of both together exceed what a visit brings back)
mkdir -p gpurun_out; L=gpurun_out/r2f.log; rm -f $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2f_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
for c in 1 2 4 5; do
timeout 300 python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/r2f_cfg${c}.json 2> gpurun_out/r2f_cfg${c}.err
tail -2 gpurun_out/r2f_cfg${c}.err >> $L
python - gpurun_out/r2f_cfg${c}.json >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g"%d["value"], "kernel ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cat $L; cat gpurun_out/r2f_bench.json; cat gpurun_out/r2f_bench_reference.json
