#!/bin/bash
# round 2, visit CC (1 GPU): span reverse kernel with padded mask arrays and skewed vol rows -- parity, shard timing, wavefronts
mkdir -p gpurun_out; L=gpurun_out/r2cc.log; rm -f $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
for N in 131072; do
  echo "== N=$N" >> $L
  CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py $N 30 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|rev end|rev live|compaction|fwd end" >> $L
  timeout 300 python scripts/prof_config3.py $N 30 aad 2>&1 | tail -4 | grep -E "step ms|kernel avg" >> $L
done
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:dupire_reverse_span" -s 1 -c 1 -f -o gpurun_out/r2cc_span python scripts/prof_config3.py 131072 3 aad > gpurun_out/r2cc_span.log 2>&1
NCU_KERNEL=dupire_reverse_span python tools/ncu_smem.py gpurun_out/r2cc_span.ncu-rep 12 > gpurun_out/r2cc_span.smem.txt 2>&1
python tools/ncu_multi.py gpurun_out/r2cc_span.ncu-rep 2>/dev/null | head -14 >> $L
rm -f gpurun_out/r2cc_span.ncu-rep
cat $L; cat gpurun_out/r2cc_span.smem.txt | cut -c1-170
