#!/bin/bash
# ncu full capture of the Dupire kernels for forward variants given as arguments (value mode = forward only without history)
mkdir -p gpurun_out
for V in "$@"; do
  CF_DUPIRE_FWD=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:dupire_ -s 3 -c 2 -o gpurun_out/prof_fwd$V -f python scripts/prof_config3.py 1048576 3 aad > gpurun_out/prof_run_fwd$V.log 2>&1
  tail -2 gpurun_out/prof_run_fwd$V.log
done
