#!/bin/bash
# round 2, visit A: parity of the new small-shard kernels (forward chunk 8, quad reverse, PDL) + timing sweep
mkdir -p gpurun_out
echo "== pytest default" > gpurun_out/r2a.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> gpurun_out/r2a.log
echo "== pytest CF_DUPIRE_REV=quad CF_DUPIRE_FWD_P=1" >> gpurun_out/r2a.log
CF_DUPIRE_REV=quad CF_DUPIRE_FWD_P=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 >> gpurun_out/r2a.log
echo "== pytest CF_DUPIRE_REV=quad (P auto)" >> gpurun_out/r2a.log
CF_DUPIRE_REV=quad timeout 900 python -m pytest tests -m gpu -x -q -k "dupire or parity or full" 2>&1 | tail -6 >> gpurun_out/r2a.log
for N in 131072 262144 1048576; do
  for V in "CF_DUPIRE_REV=classic CF_DUPIRE_FWD_CH=4" "CF_DUPIRE_REV=classic CF_DUPIRE_FWD_CH=8" "CF_DUPIRE_REV=quad CF_DUPIRE_FWD_CH=8 CF_PDL=0" "CF_DUPIRE_REV=quad CF_DUPIRE_FWD_CH=8 CF_PDL=1"; do
    echo "== N=$N $V" >> gpurun_out/r2a.log
    env $V timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -2 >> gpurun_out/r2a.log
  done
  echo "== N=$N quad CH=8 PDL=1 own stream" >> gpurun_out/r2a.log
  CF_PROF_STREAM=1 CF_DUPIRE_REV=quad timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -2 >> gpurun_out/r2a.log
  echo "== N=$N value CH=4 / CH=8" >> gpurun_out/r2a.log
  CF_DUPIRE_FWD_CH=4 timeout 300 python scripts/prof_config3.py $N 20 value 2>&1 | tail -1 >> gpurun_out/r2a.log
  CF_DUPIRE_FWD_CH=8 timeout 300 python scripts/prof_config3.py $N 20 value 2>&1 | tail -1 >> gpurun_out/r2a.log
done
# per-kernel durations (ncu launch list) at 2^17 paths: classic vs new
for V in "CF_DUPIRE_REV=classic CF_DUPIRE_FWD_CH=4" "CF_DUPIRE_REV=quad CF_DUPIRE_FWD_CH=8"; do
  tag=$(echo $V | tr ' =' '__')
  env $V timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/r2a_launches_$tag.csv python scripts/prof_config3.py 131072 8 aad > /dev/null 2>&1
done
cat gpurun_out/r2a.log
