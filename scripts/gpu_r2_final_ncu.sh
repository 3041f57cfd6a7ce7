#!/bin/bash
# round 2, final 1-GPU visit, part 2: ncu --set full captures of the hot kernels (config 3 at 2^20 and at a 2^17 shard,
# configs 2, 4, 5), summarised on the box (tools/ncu_multi.py, tools/ncu_lines.py); only the two largest reports come back
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c $c -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  python tools/ncu_multi.py gpurun_out/$name.ncu-rep > gpurun_out/$name.summary.txt 2>&1
  python tools/ncu_lines.py gpurun_out/$name.ncu-rep 30 > gpurun_out/$name.lines.txt 2>&1
}
cap r2f_dupire_aad dupire_ 3 3 python scripts/prof_config3.py 1048576 3 aad
cap r2f_dupire_shard dupire_ 3 3 python scripts/prof_config3.py 131072 3 aad
cap r2f_dlm dlm_kernel 1 1 python scripts/prof_configs.py 5 1048576
cap r2f_bs "dupire_forward4|bs_reverse" 2 2 python scripts/prof_configs.py 2 1048576
cap r2f_multi dupire_europeans_multi 0 1 python -c "
import sys; sys.path.insert(0,'.')
import bench
from compfinance_b200.api import CompFinance
cf=CompFinance(device=0); m=bench._put_config(cf,4)
for _ in range(2): cf.aad_risk_multi(m,'bench_prd',1<<20,sobol=False)
"
rm -f gpurun_out/r2f_dupire_shard.ncu-rep gpurun_out/r2f_bs.ncu-rep gpurun_out/r2f_multi.ncu-rep
ls -la gpurun_out | tail -30
cat gpurun_out/r2f_dupire_aad.summary.txt | head -60
