#!/bin/bash
# Dev tool (under gpurun, one GPU): bench.py over a list of configurations WL:MASK:LANES:MB[:LIB]
#   MASK = WSO_WARP_CORE, LANES = WSO_LANES, MB = WSO_W_BUDGET_MB, LIB = variant name in build/variants (optional)
# usage: bash tools/gpu_matrix.sh TAG [--test] cfg...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$1" = "--test" ]; then shift
  WSO_WARP_CORE=7 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -n 3 $OUT/pytest_gpu.log
fi
for cfg in "$@"; do
  IFS=: read wl m lanes mb lib <<< "$cfg"
  name=${wl}_m${m}_l${lanes}_mb${mb}${lib:+_$lib}
  libenv=""; [ -n "$lib" ] && libenv="WSO_LIB_PATH=$PWD/build/variants/libwsocean_$lib.so"
  env $libenv WSO_WARP_CORE=$m WSO_LANES=$lanes WSO_W_BUDGET_MB=$mb timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err
done
python tools/summ.py $OUT/bench_*.json
