#!/bin/bash
# round 2, visit J (1 GPU): span 8 warps + balanced shares -- parity + timing; ncu of the DLM and BS kernels for planning
mkdir -p gpurun_out; L=gpurun_out/r2j.log; rm -f $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $L
for N in 131072 196608 262144; do
  for V in "CF_DUPIRE_REV=classic" "CF_DUPIRE_REV=span"; do
    echo "== N=$N $V" >> $L
    env $V CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|rev live|rev end" >> $L
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dlm_kernel -s 1 -c 1 -o gpurun_out/r2j_dlm python scripts/prof_configs.py 5 1048576 > gpurun_out/r2j_ncu_dlm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -o gpurun_out/r2j_bs python scripts/prof_configs.py 2 1048576 > gpurun_out/r2j_ncu_bs.log 2>&1
cat $L
