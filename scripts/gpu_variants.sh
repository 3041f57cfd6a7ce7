#!/bin/bash
# GPU visit: parity tests with the default forward variant, then kernel timing for each variant given as argument
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for V in "$@"; do
  for mode in aad value; do
    CF_DUPIRE_FWD=$V timeout 300 python scripts/prof_config3.py 1048576 8 $mode 2>&1 | tail -1 | sed "s/^/FWD=$V /"
  done
done | tee gpurun_out/variants.log
