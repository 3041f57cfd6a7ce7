#!/bin/bash
# GPU-box visit: parity tests, bench (1 GPU), e2e breakdown, secondary configs.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
CF_TIMING=1 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
timeout 900 python tools/bench_configs.py --reps 3 > gpurun_out/configs.json 2> gpurun_out/configs.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -8 gpurun_out/e2e_breakdown.log; cat gpurun_out/configs.json; tail -5 gpurun_out/configs.err
