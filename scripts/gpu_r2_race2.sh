#!/bin/bash
# round 2: racecheck of the span reverse kernel after skipping weight-0 targets; then the full suite and the shard timing
mkdir -p gpurun_out; L=gpurun_out/r2race2.log; rm -f $L
CF_DUPIRE_REV=span timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 8 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_config3_dupire_barrier_golden or test_step_counts_around_the_chunk_size_vs_reference" > gpurun_out/r2race2_span.txt 2>&1
grep -E "Error|Warning|RACECHECK SUMMARY|passed|failed" gpurun_out/r2race2_span.txt | cut -c1-200 | head -8 >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
CF_DEBUG_TIMES=1 timeout 300 python scripts/prof_config3.py 131072 30 aad 2>&1 | tail -18 | grep -E "step ms|kernel avg|rev sweep|compaction" >> $L
cat $L
