#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dlm_kernel -s 1 -c 1 -f -o gpurun_out/prof_dlm5 python scripts/prof_configs.py 5 1048576 > gpurun_out/prof_dlm5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/prof_bs2 python scripts/prof_configs.py 2 1048576 > gpurun_out/prof_bs2.log 2>&1
python scripts/prof_configs.py 5v 4194304
tail -3 gpurun_out/prof_dlm5.log gpurun_out/prof_bs2.log
