#!/bin/bash
# round 2: compute-sanitizer (memcheck, then racecheck on shared memory) over small cases of every kernel family
# changed this round: span reverse, Black-Scholes fast path, displaced-model kernel v2, K-multi, call ladders
mkdir -p gpurun_out; L=gpurun_out/r2san.log; rm -f $L
SEL="test_config2_bs_barrier or test_config3_dupire_barrier_golden or test_step_counts_around_the_chunk_size_vs_reference or test_every_asset_count_bucket_vs_reference or test_config4_itemised_risk_matrix or test_arbitrary_shard_boundaries_add_up_displaced_model or test_baskets_vs_reference or test_black_scholes_europeans"
for TOOL in memcheck racecheck; do
  echo "== compute-sanitizer --tool $TOOL" >> $L
  timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 99 --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r2san_$TOOL.txt 2>&1
  echo "exit $?" >> $L
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" gpurun_out/r2san_$TOOL.txt | sort | uniq -c | sort -rn | head -12 >> $L
done
cat $L
