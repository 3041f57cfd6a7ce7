#!/bin/bash
# full GPU suite + the secondary configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/bench_configs.py --reps 3 > gpurun_out/configs.json 2> gpurun_out/configs.err
cat gpurun_out/configs.json | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['config'], '%.3f ms' % r['ms_per_call'], '%.3g paths/s' % r['paths_per_sec_e2e'], 'x%.0f' % r.get('speedup_vs_cpu_reference', 0))"
tail -3 gpurun_out/configs.err
