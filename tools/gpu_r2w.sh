#!/bin/bash
# Dev tool (under gpurun, one GPU): W budget per chunk (= tile-frames per launch = waves per kernel) on every workload.  usage: gpu_r2w.sh TAG "BUDGETS_MB"
TAG=${1:-r2w}; MBS=${2:-"64 140"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for wl in c2 c3 c4; do for mb in $MBS; do
  WSO_WARP_CORE=$([ $wl = c2 ] && echo 2 || echo 0) WSO_W_BUDGET_MB=$mb timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_mb$mb.json 2> $OUT/bench_${wl}_mb$mb.err
done; done
python tools/summ.py $OUT/bench_*.json
