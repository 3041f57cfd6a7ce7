#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dlm_kernel -s 1 -c 1 -f -o gpurun_out/prof_dlm5 python scripts/prof_configs.py 5 1048576 > gpurun_out/prof_dlm5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dlm_kernel -s 1 -c 1 -f -o gpurun_out/prof_dlm5v python scripts/prof_configs.py 5v 1048576 > gpurun_out/prof_dlm5v.log 2>&1
