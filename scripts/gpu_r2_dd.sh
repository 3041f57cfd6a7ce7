#!/bin/bash
# round 2, visit DD (1 GPU): shared-memory wavefronts per source line of the forward kernel (2^20 paths) and the classic reverse
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:dupire_forward4|dupire_reverse_kernel" -s 2 -c 2 -f -o gpurun_out/r2dd python scripts/prof_config3.py 1048576 3 aad > gpurun_out/r2dd.log 2>&1
NCU_KERNEL=dupire_forward4 python tools/ncu_smem.py gpurun_out/r2dd.ncu-rep 22 > gpurun_out/r2dd_fwd.smem.txt 2>&1
NCU_KERNEL=dupire_reverse_kernel python tools/ncu_smem.py gpurun_out/r2dd.ncu-rep 16 > gpurun_out/r2dd_rev.smem.txt 2>&1
rm -f gpurun_out/r2dd.ncu-rep
cat gpurun_out/r2dd_fwd.smem.txt | cut -c1-190; cat gpurun_out/r2dd_rev.smem.txt | cut -c1-190
