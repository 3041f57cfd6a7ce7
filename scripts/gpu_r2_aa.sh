#!/bin/bash
# round 2, visit AA (1 GPU): shard-boundary tests through cfx_run_range, then the whole GPU suite
mkdir -p gpurun_out; L=gpurun_out/r2aa.log; rm -f $L
timeout 900 python -m pytest tests -m gpu -x -q -k "shard_boundaries" 2>&1 | tail -15 >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $L
cat $L
