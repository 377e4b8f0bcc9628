#!/bin/bash
# Dev tool (under gpurun, one GPU): GPU parity suite, one c2 bench of the in-tree build, then the variant libraries.
TAG=${1:-r1r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 60 python bench.py --workload c2 --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench c2 rc=$?"
for so in build/variants/libwsocean_n*.so; do
  [ -f $so ] || continue
  name=$(basename $so .so); name=${name#libwsocean_}
  case $name in n9_*) wl=c4;; n10_*) wl=c2;; n11_*) wl=c3;; *) wl=c2;; esac
  WSO_LIB_PATH=$PWD/$so timeout 40 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/var_${name}.json 2> $OUT/var_${name}.err
done
tail -n 12 $OUT/pytest_gpu.log; python tools/summ.py $OUT/bench_c?.json $OUT/var_*.json
