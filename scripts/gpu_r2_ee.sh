#!/bin/bash
# round 2, visit EE (1 GPU): where a first call (no resident session) spends its time
mkdir -p gpurun_out
CF_HOST_CACHE=0 CF_TIMING=1 timeout 300 python scripts/e2e_breakdown.py 2>&1 | tail -24
