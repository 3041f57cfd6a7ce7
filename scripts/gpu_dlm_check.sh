#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_xl.py tests/test_dist.py -m gpu -x -q > gpurun_out/pytest_dlm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dlm.log
tail -15 gpurun_out/pytest_dlm.log
timeout 300 python scripts/prof_configs.py 5 4194304
timeout 300 python scripts/prof_configs.py 5v 4194304
