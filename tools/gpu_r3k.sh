#!/bin/bash
# Dev tool (under gpurun, one GPU): L2 residency experiments for the intermediate W - persisting access-policy window (env) and
# evict-last store hint (variant build) - plus any other build/variants/libwsocean_all_*.so, on C2 / C3 / C4.   usage: gpu_r3k.sh TAG
TAG=${1:-r3k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for wl in c2 c3 c4; do
  timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_new.json 2> $OUT/bench_${wl}_new.err
  for mb in 48 96; do
    WSO_EXP_L2_PERSIST_MB=$mb timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_persist$mb.json 2> $OUT/bench_${wl}_persist$mb.err
  done
  WSO_EXP_L2_PERSIST_MB=96 WSO_EXP_L2_PERSIST_RATIO=0.6 timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_persist96r60.json 2> $OUT/bench_${wl}_persist96r60.err
  for so in build/variants/libwsocean_all_*.so; do
    [ -f $so ] || continue
    name=$(basename $so .so); name=${name#libwsocean_all_}
    WSO_LIB_PATH=$PWD/$so timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_$name.json 2> $OUT/bench_${wl}_$name.err
  done
done
python tools/summ.py $OUT/bench_*.json
grep -h "persisting L2" $OUT/*.err | sort | uniq -c
