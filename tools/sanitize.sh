#!/bin/bash
# compute-sanitizer over the hot-path kernels (under gpurun, one GPU).   usage: bash tools/sanitize.sh [TAG]
#   memcheck  : out-of-bounds / misaligned global, shared and local accesses
#   racecheck : shared-memory hazards between the threads of a CTA - the per-line named barriers, the __syncwarp-only
#               exchanges of the warp-per-line kernels, the line buffers reused across the bulk-copy pipeline, the
#               distributed-shared-memory reads of the cluster-pair K2
#   synccheck : divergent / mismatched barriers (named barriers with the wrong participant count, shuffles)
# Cases: the parity suite at N = 16..512 on both kernel sets (WSO_WARP_CORE=0 / 7), the batched 512^2 tilings, and one
# slab case with the two-CTA cluster kernel.  The sanitizer slows a launch 10-100x: sizes stay small, logs are summaries.
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG
mkdir -p $OUT
# SAN_TOOLS / SAN_CORES / SAN_SEL / SAN_SLAB narrow a run (e.g. a re-check after a kernel change with a small GPU budget)
SEL='(test_reference_fixtures or test_one_hot_layout or test_batch_matches_single_frames or test_independent_tiles_in_one_batch or (test_all_tile_sizes_vs_oracle and (16 or 32 or 64 or 128 or 256 or 512)) or (test_bulk_tilings_vs_oracle and 512) or (test_jacobian_channel and 64))'
SLAB='test_slab_world1_fixtures'
SEL=${SAN_SEL:-$SEL}; SLAB=${SAN_SLAB:-$SLAB}
rc_all=0
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  for core in ${SAN_CORES:-0 7}; do
    log=$OUT/${tool}_core$core.log
    WSO_WARP_CORE=$core timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
      python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" > $log 2>&1
    rc=$?; [ $rc -ne 0 ] && rc_all=1
    echo "$tool core=$core rc=$rc: $(grep -E 'passed|failed|error' $log | tail -n 1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -n 1)"
  done
  [ "$SLAB" = none ] && continue
  log=$OUT/${tool}_slab.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests/test_slab.py -m gpu -x -q -k "$SLAB" > $log 2>&1
  rc=$?; [ $rc -ne 0 ] && rc_all=1
  echo "$tool slab rc=$rc: $(grep -E 'passed|failed|error' $log | tail -n 1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -n 1)"
done | tee $OUT/summary.txt
exit $rc_all
