#!/bin/bash
# round 2, visit B: phase timestamps of the small-shard kernels + ncu full capture of the quad reverse kernel
mkdir -p gpurun_out
for V in "CF_DUPIRE_REV=quad CF_PDL=0" "CF_DUPIRE_REV=quad CF_PDL=1"; do
  echo "== N=131072 $V" >> gpurun_out/r2b.log
  env $V CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py 131072 12 aad 2>&1 | tail -18 >> gpurun_out/r2b.log
done
echo "== N=1048576 quad" >> gpurun_out/r2b.log
CF_DUPIRE_REV=quad CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py 1048576 12 aad 2>&1 | tail -18 >> gpurun_out/r2b.log
CF_DUPIRE_REV=quad timeout 900 ncu --set full --clock-control none --import-source on -k regex:dupire_reverse_quad -s 4 -c 1 -o gpurun_out/r2b_revq_small python scripts/prof_config3.py 131072 6 aad > gpurun_out/r2b_ncu.log 2>&1
CF_DUPIRE_FWD_CH=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dupire_forward4 -s 4 -c 1 -o gpurun_out/r2b_fwd8_small python scripts/prof_config3.py 131072 6 aad >> gpurun_out/r2b_ncu.log 2>&1
cat gpurun_out/r2b.log
