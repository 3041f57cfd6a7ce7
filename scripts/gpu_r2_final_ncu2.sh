#!/bin/bash
# round 2, final captures of the kernels changed after gpu_r2_final_ncu.sh: K-multi with the call-ladder sums,
# the Black-Scholes reverse kernel on 24 warps; summarised on the box
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c $c -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  python tools/ncu_multi.py gpurun_out/$name.ncu-rep > gpurun_out/$name.summary.txt 2>&1
  python tools/ncu_lines.py gpurun_out/$name.ncu-rep 30 > gpurun_out/$name.lines.txt 2>&1
}
cap r2g_bs "dupire_forward4|bs_reverse" 2 2 python scripts/prof_configs.py 2 1048576
cap r2g_multi dupire_europeans_multi 0 1 python -c "
import sys; sys.path.insert(0,'.')
import bench
from compfinance_b200.api import CompFinance
cf=CompFinance(device=0); m=bench._put_config(cf,4)
for _ in range(2): cf.aad_risk_multi(m,'bench_prd',1<<20,sobol=False)
"
rm -f gpurun_out/r2g_bs.ncu-rep gpurun_out/r2g_multi.ncu-rep
grep -v stall gpurun_out/r2g_multi.summary.txt | head -16; grep -v stall gpurun_out/r2g_bs.summary.txt | tail -16
