#!/bin/bash
# round 2: per-line stall reasons (tools/ncu_stalls.py) of the north-star kernels at 2^20 paths and at a 2^17-path shard,
# and of the displaced-model kernel
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c $c -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  NCU_KERNEL=$k python tools/ncu_stalls.py gpurun_out/$name.ncu-rep 28 > gpurun_out/$name.stalls.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap r2s_fwd dupire_forward4 1 1 python scripts/prof_config3.py 1048576 3 aad
cap r2s_rev "dupire_reverse_kernel" 1 1 python scripts/prof_config3.py 1048576 3 aad
cap r2s_span "dupire_reverse_span" 1 1 python scripts/prof_config3.py 131072 3 aad
cap r2s_fwd1 dupire_forward4 1 1 python scripts/prof_config3.py 131072 3 aad
cap r2s_dlm dlm_kernel 1 1 python scripts/prof_configs.py 5 1048576
for n in fwd rev span fwd1 dlm; do echo "=== $n"; head -22 gpurun_out/r2s_$n.stalls.txt | cut -c1-190; done
