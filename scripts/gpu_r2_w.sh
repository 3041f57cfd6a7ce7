#!/bin/bash
# round 2, visit W (1 GPU): where K-multi waits for instructions (per-line stall reasons)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:dupire_europeans_multi" -s 0 -c 1 -f -o gpurun_out/r2w_multi python -c "
import sys; sys.path.insert(0,'.')
import bench
from compfinance_b200.api import CompFinance
cf=CompFinance(device=0); m=bench._put_config(cf,4)
for _ in range(2): cf.aad_risk_multi(m,'bench_prd',1<<20,sobol=False)
" > gpurun_out/r2w_multi.log 2>&1
python tools/ncu_stalls.py gpurun_out/r2w_multi.ncu-rep 40 > gpurun_out/r2w_multi.stalls.txt 2>&1
ncu -i gpurun_out/r2w_multi.ncu-rep --page details 2>/dev/null | grep -iE "instruction|icache|branch|divergen|Avg. Active Threads|Local" | head -30 > gpurun_out/r2w_multi.details.txt
rm -f gpurun_out/r2w_multi.ncu-rep
cat gpurun_out/r2w_multi.stalls.txt | cut -c1-200; cat gpurun_out/r2w_multi.details.txt
