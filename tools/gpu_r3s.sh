#!/bin/bash
# Dev tool (under gpurun, one GPU): tiling sweep with the final kernels - build/variants/libwsocean_n{9,10,11}_*.so on C4 / C2 / C3.   usage: gpu_r3s.sh TAG
TAG=${1:-r3s}; OUT=gpurun_out/$TAG; mkdir -p $OUT
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for pair in 10:c2 9:c4 11:c3; do
  logn=${pair%%:*}; wl=${pair##*:}
  for so in build/variants/libwsocean_n${logn}_*.so; do
    name=$(basename $so .so); name=${name#libwsocean_}
    WSO_LIB_PATH=$PWD/$so timeout 120 python bench.py --workload $wl $B > $OUT/bench_${wl}_$name.json 2> $OUT/bench_${wl}_$name.err
  done
done
python tools/summ.py $OUT/bench_*.json
