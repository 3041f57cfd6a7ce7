#!/bin/bash
# quick GPU visit: parity tests, bench, kernel timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for mode in aad value; do timeout 300 python scripts/prof_config3.py 1048576 6 $mode 2>&1 | tail -1; done | tee gpurun_out/kern.log
