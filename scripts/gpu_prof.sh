#!/bin/bash
# ncu full capture of the Dupire AAD kernel for the given paths-per-thread values
mkdir -p gpurun_out
for P in "$@"; do
  CF_DUPIRE_P=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:dupire_kernel -s 1 -c 1 -o gpurun_out/prof_v3_P$P -f python scripts/prof_config3.py 1048576 3 aad > gpurun_out/prof_run_P$P.log 2>&1
  tail -2 gpurun_out/prof_run_P$P.log
done
