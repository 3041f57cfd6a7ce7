#!/bin/bash
# round 2, visit C: span reverse kernel -- parity + timing
mkdir -p gpurun_out; L=gpurun_out/r2c.log; rm -f $L
echo "== pytest default (span)" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 >> $L
echo "== pytest FWD_P=1" >> $L
CF_DUPIRE_FWD_P=1 timeout 900 python -m pytest tests -m gpu -x -q -k "dupire or parity or full" 2>&1 | tail -6 >> $L
for N in 131072 1048576; do
  for V in "CF_DUPIRE_REV=classic" "CF_DUPIRE_REV=span CF_PDL=0" "CF_DUPIRE_REV=span CF_PDL=1"; do
    echo "== N=$N $V" >> $L
    env $V CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -18 >> $L
  done
done
for N in 262144 524288; do
  for V in "CF_DUPIRE_REV=classic" "CF_DUPIRE_REV=span"; do
    echo "== N=$N $V" >> $L
    env $V timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -2 >> $L
  done
done
CF_DUPIRE_REV=span timeout 900 ncu --set full --clock-control none --import-source on -k regex:dupire_reverse_span -s 4 -c 1 -o gpurun_out/r2c_revs_small python scripts/prof_config3.py 131072 6 aad > gpurun_out/r2c_ncu.log 2>&1
CF_DUPIRE_REV=span timeout 900 ncu --set full --clock-control none --import-source on -k regex:dupire_reverse_span -s 4 -c 1 -o gpurun_out/r2c_revs_full python scripts/prof_config3.py 1048576 6 aad >> gpurun_out/r2c_ncu.log 2>&1
cat $L
