#!/bin/bash
# Dev tool (under gpurun, one GPU): A/B of the variant libraries build/variants/libwsocean_n<LOGN>_*.so over a list of kernel masks.
# usage: gpu_r2v.sh TAG "MASKS" [WORKLOAD]
TAG=${1:-r2v}; MASKS=${2:-"2 18 66"}; WL=${3:-c2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for so in build/variants/libwsocean_n*.so; do
  [ -f $so ] || continue
  name=$(basename $so .so); name=${name#libwsocean_}
  for m in $MASKS; do
    WSO_WARP_CORE=$m WSO_LIB_PATH=$PWD/$so timeout 100 python bench.py --workload $WL $B > $OUT/var_${name}_m$m.json 2> $OUT/var_${name}_m$m.err
  done
done
python tools/summ.py $OUT/var_*.json
