#!/bin/bash
# kernel timing at a given path count for forward launch shapes; needs a CF_SWEEP=1 build
N=$1; shift
for V in "$@"; do
  for mode in aad value; do
    CF_DUPIRE_FWD=$V timeout 300 python scripts/prof_config3.py $N 12 $mode 2>&1 | tail -1 | sed "s/^/N=$N FWD=$V /"
  done
done | tee -a gpurun_out/sweep_n.log
