#!/bin/bash
# round 2, visit I (1 GPU): span kernel warps-per-block variants (register budget vs thread-level parallelism) + new parity tests
mkdir -p gpurun_out; L=gpurun_out/r2i.log; rm -f $L
timeout 600 python -m pytest tests -m gpu -x -q -k "superbucket or bump or uniform" 2>&1 | tail -6 >> $L
for N in 131072 262144; do
  for W in 8 12 16; do
    echo "== N=$N span warps=$W" >> $L
    LIBV=""; if [ "$W" != "16" ]; then LIBV="CF_B200_LIB=/root/repo/compfinance_b200/lib/exp/libcf_b200_w$W.so"; fi
    env $LIBV CF_DUPIRE_REV=span CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 20 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|rev end" >> $L
  done
done
cat $L
