#!/bin/bash
# Dev tool (under gpurun, one GPU): column-pair barrier behind K1's fused front end - parity (short timeouts: a wrong barrier count
# hangs), racecheck / synccheck on the fused tilings, A/B against build/variants/libwsocean_all_nocolpair.so.   usage: gpu_r3t.sh TAG
TAG=${1:-r3t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 120 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "(test_all_tile_sizes_vs_oracle and (512 or 1024)) or (test_bulk_tilings_vs_oracle and (512 or 1024))" > $OUT/pytest_quick.log 2>&1; rc=$?; echo "quick parity rc=$rc"; tail -3 $OUT/pytest_quick.log
[ $rc -ne 0 ] && exit 1
for tool in racecheck synccheck; do
  timeout 200 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 10 python -m pytest tests/test_parity_gpu.py -m gpu -x -q \
    -k "(test_all_tile_sizes_vs_oracle and 1024) or (test_bulk_tilings_vs_oracle and 512)" > $OUT/$tool.log 2>&1
  echo "$tool rc=$?: $(grep -E 'passed|failed' $OUT/$tool.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.log | tail -1)"
done
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for rep in 1 2; do
  for wl in c2 c4; do
    timeout 100 python bench.py --workload $wl $B > $OUT/bench_${wl}_new$rep.json 2> $OUT/bench_${wl}_new$rep.err
    WSO_LIB_PATH=$PWD/build/variants/libwsocean_all_nocolpair.so timeout 100 python bench.py --workload $wl $B > $OUT/bench_${wl}_nocolpair$rep.json 2> $OUT/bench_${wl}_nocolpair$rep.err
  done
done
python tools/summ.py $OUT/bench_*.json
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
