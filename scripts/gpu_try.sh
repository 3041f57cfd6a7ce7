#!/bin/bash
# quick GPU visit: parity tests, then kernel timing for each variant (CF_DUPIRE_P values given as arguments)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for P in "$@"; do
  for mode in aad value; do
    CF_DUPIRE_P=$P timeout 300 python scripts/prof_config3.py 1048576 6 $mode 2>&1 | tail -1 | sed "s/^/P=$P /"
  done
done | tee gpurun_out/psweep.log
