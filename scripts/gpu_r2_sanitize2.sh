#!/bin/bash
# round 2: compute-sanitizer memcheck over the two-device context (peer exchange inside the reduction kernels)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 10 python tests/multidev_check.py > gpurun_out/r2san2_multidev.txt 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|OK|ok|passed|FAIL|Invalid" gpurun_out/r2san2_multidev.txt | tail -8
