#!/bin/bash
# ncu full capture of the Dupire AAD kernels (forward + reverse) for the given CF_DUPIRE_P values
mkdir -p gpurun_out
for P in "$@"; do
  CF_DUPIRE_P=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:dupire_ -s 3 -c 2 -o gpurun_out/prof_v3_P$P -f python scripts/prof_config3.py 1048576 3 aad > gpurun_out/prof_run_P$P.log 2>&1
  tail -2 gpurun_out/prof_run_P$P.log
done
