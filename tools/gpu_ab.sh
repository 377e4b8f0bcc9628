#!/bin/bash
# Dev tool (under gpurun): GPU parity tests on the in-tree build, bench of every workload, A/B of the variant
# libraries in build/variants/ on one workload, one ncu --set full capture with source.
# usage: bash tools/gpu_ab.sh TAG [LOGN WORKLOAD]
TAG=${1:-ab}; LOGN=${2:-10}; WL=${3:-c2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
for wl in c2 c3 c4 c1; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"
done
for so in build/variants/libwsocean_n${LOGN}_*.so; do
  name=$(basename $so .so); name=${name#libwsocean_}
  WSO_LIB_PATH=$PWD/$so timeout 120 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
done
for n in 512 1024; do timeout 120 tools/lat_bench $n 2000 >> $OUT/lat_bench.jsonl 2>> $OUT/lat_bench.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wso_ -s 30 -c 3 -o $OUT/prof_c2 \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_c2.log 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/lat_bench.jsonl; python tools/summ.py $OUT/bench_c?.json $OUT/n${LOGN}_*.json
