#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for lib in default build/libcf_w20.so build/libcf_w16.so; do
  for mode in aad value; do
    if [ $lib = default ]; then CF_DUPIRE_P=4 timeout 300 python scripts/prof_config3.py 1048576 6 $mode 2>&1 | tail -1 | sed "s|^|$lib |";
    else CF_B200_LIB=$PWD/$lib CF_DUPIRE_P=4 timeout 300 python scripts/prof_config3.py 1048576 6 $mode 2>&1 | tail -1 | sed "s|^|$lib |"; fi
  done
done | tee gpurun_out/wsweep.log
