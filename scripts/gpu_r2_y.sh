#!/bin/bash
# round 2, visit Y (1 GPU): generic kernel on config 4 (value and aggregate risk), K-multi after the fix: stall reasons
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count, python snippet
  local name=$1 k=$2 s=$3 c=$4 code=$5
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c $c -f -o gpurun_out/$name python -c "$code" > gpurun_out/$name.log 2>&1
  python tools/ncu_multi.py gpurun_out/$name.ncu-rep > gpurun_out/$name.summary.txt 2>&1
  python tools/ncu_stalls.py gpurun_out/$name.ncu-rep 25 > gpurun_out/$name.stalls.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
PRE="import sys; sys.path.insert(0,'.'); import bench, numpy as np
from compfinance_b200.api import CompFinance
cf=CompFinance(device=0); m=bench._put_config(cf,4)"
cap r2y_value path_kernel 1 1 "$PRE
for _ in range(3): cf.value(m,'bench_prd',1<<20,sobol=False)"
cap r2y_aggr path_kernel 1 1 "$PRE
w=0.5+np.cos(np.arange(720))
for _ in range(3): cf.aad_risk_aggregate(m,'bench_prd',w,1<<20,sobol=False)"
cap r2y_multi dupire_europeans_multi 1 1 "$PRE
for _ in range(3): cf.aad_risk_multi(m,'bench_prd',1<<20,sobol=False)"
for n in value aggr multi; do echo "=== $n"; grep -v "stall" gpurun_out/r2y_$n.summary.txt | head -9; head -30 gpurun_out/r2y_$n.stalls.txt | cut -c1-190; done
