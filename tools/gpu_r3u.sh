#!/bin/bash
# Dev tool (under gpurun, one GPU): split + store dealt to the field groups (barrier per field group) - quick parity with a short timeout,
# racecheck, A/B against build/variants/libwsocean_all_nofieldsplit.so on C2 / C3 / C4, then the GPU suite.   usage: gpu_r3u.sh TAG
TAG=${1:-r3u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 150 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "(test_all_tile_sizes_vs_oracle and (256 or 512 or 1024 or 2048)) or test_bulk_tilings_vs_oracle" > $OUT/pytest_quick.log 2>&1; rc=$?; echo "quick parity rc=$rc"; tail -3 $OUT/pytest_quick.log
[ $rc -ne 0 ] && exit 1
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 77 --print-limit 10 python -m pytest tests/test_parity_gpu.py -m gpu -x -q \
  -k "(test_all_tile_sizes_vs_oracle and (256 or 1024)) or (test_bulk_tilings_vs_oracle and 512)" > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?: $(grep -E 'passed|failed' $OUT/racecheck.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' $OUT/racecheck.log | tail -1)"
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for wl in c2 c3 c4; do
  timeout 100 python bench.py --workload $wl $B > $OUT/bench_${wl}_new.json 2> $OUT/bench_${wl}_new.err
  WSO_LIB_PATH=$PWD/build/variants/libwsocean_all_nofieldsplit.so timeout 100 python bench.py --workload $wl $B > $OUT/bench_${wl}_nofieldsplit.json 2> $OUT/bench_${wl}_nofieldsplit.err
done
python tools/summ.py $OUT/bench_*.json
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
