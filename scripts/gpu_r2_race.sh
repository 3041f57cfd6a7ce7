#!/bin/bash
# round 2: racecheck on the Black-Scholes barrier test (after aligning the warp totals) and on the Dupire barrier test
# with each reverse form, to attribute the warp-level warnings
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 8 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_config2_bs_barrier" > gpurun_out/r2race_bs.txt 2>&1
for F in classic span; do
CF_DUPIRE_REV=$F timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 8 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_config3_dupire_barrier_golden" > gpurun_out/r2race_dup_$F.txt 2>&1
done
for f in bs dup_classic dup_span; do echo "=== $f"; grep -E "Error|Warning|RACECHECK SUMMARY|passed|failed" gpurun_out/r2race_$f.txt | cut -c1-200 | head -12; done
