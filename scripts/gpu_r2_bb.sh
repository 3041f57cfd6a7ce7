#!/bin/bash
# round 2, visit BB (1 GPU): shared-memory wavefronts per source line of the span reverse kernel (2^17-path shard)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:dupire_reverse_span" -s 1 -c 1 -f -o gpurun_out/r2bb_span python scripts/prof_config3.py 131072 3 aad > gpurun_out/r2bb_span.log 2>&1
NCU_KERNEL=dupire_reverse_span python tools/ncu_smem.py gpurun_out/r2bb_span.ncu-rep 30 > gpurun_out/r2bb_span.smem.txt 2>&1
rm -f gpurun_out/r2bb_span.ncu-rep
cat gpurun_out/r2bb_span.smem.txt | cut -c1-200
