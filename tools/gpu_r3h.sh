#!/bin/bash
# Dev tool (under gpurun, one GPU): GPU test suite; A/B of the in-tree library against build/variants/libwsocean_all_*.so on C2 / C3 / C4
# (default kernel choice), and the n11_* variants on C3.   usage: gpu_r3h.sh TAG
TAG=${1:-r3h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -n 3 $OUT/pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for wl in c2 c3 c4; do
  timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_new.json 2> $OUT/bench_${wl}_new.err
  for so in build/variants/libwsocean_all_*.so; do
    [ -f $so ] || continue
    name=$(basename $so .so); name=${name#libwsocean_all_}
    WSO_LIB_PATH=$PWD/$so timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_$name.json 2> $OUT/bench_${wl}_$name.err
  done
done
for so in build/variants/libwsocean_n11_*.so; do
  [ -f $so ] || continue
  name=$(basename $so .so); name=${name#libwsocean_}
  WSO_LIB_PATH=$PWD/$so timeout 200 python bench.py --workload c3 $B > $OUT/bench_c3_$name.json 2> $OUT/bench_c3_$name.err
done
python tools/summ.py $OUT/bench_*.json
