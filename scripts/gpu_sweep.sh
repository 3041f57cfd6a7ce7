#!/bin/bash
# kernel timing for forward launch shapes (CF_DUPIRE_FWD = 100 P + warps); needs a CF_SWEEP=1 build
mkdir -p gpurun_out
for V in "$@"; do
  for mode in aad value; do
    CF_DUPIRE_FWD=$V timeout 300 python scripts/prof_config3.py 1048576 8 $mode 2>&1 | tail -1 | sed "s/^/FWD=$V /"
  done
done | tee gpurun_out/sweep.log
