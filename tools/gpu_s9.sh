#!/bin/bash
# Dev tool (under gpurun, one GPU): parity tests on the in-tree build, bench c1-c4, then every variant library in
# build/variants/ (name n<LOGN>_*) on the workload of its size.   usage: bash tools/gpu_s9.sh TAG
TAG=${1:-r1i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
[ -x tools/lat_bench ] || make -C examples lat_bench > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 420 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
for wl in c2 c3 c4 c1; do
  timeout 200 python bench.py --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"
done
for n in 512 1024; do timeout 60 tools/lat_bench $n 2000 >> $OUT/lat_bench.jsonl 2>> $OUT/lat_bench.err; done
for so in build/variants/libwsocean_n*.so; do
  [ -f $so ] || continue
  name=$(basename $so .so); name=${name#libwsocean_}
  case $name in n9_lat*) wl=c1;; n9_*) wl=c4;; n10_*) wl=c2;; n11_*) wl=c3;; *) wl=c2;; esac
  case $name in *lat*)  # single-tile latency through the C ABI with the variant library (RUNPATH yields to LD_LIBRARY_PATH)
    mkdir -p /tmp/v_$name; ln -sf $PWD/$so /tmp/v_$name/libwsocean.so
    echo "$name $(LD_LIBRARY_PATH=/tmp/v_$name timeout 60 tools/lat_bench 512 2000)" >> $OUT/lat_bench_variants.txt 2>> $OUT/lat_bench.err;;
  esac
  WSO_LIB_PATH=$PWD/$so timeout 100 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/var_${name}.json 2> $OUT/var_${name}.err
done
# W scratch budget per chunk (WSO_W_BUDGET_MB, default 48): tile-frames per launch triple
if [ "$2" = budget ]; then
  for cfg in c2:32 c2:48 c2:96 c3:128 c4:32 c4:48 c4:128; do
    wl=${cfg%%:*}; mb=${cfg##*:}
    WSO_W_BUDGET_MB=$mb timeout 100 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/var_budget_${wl}_${mb}mb.json 2> $OUT/var_budget_${wl}_${mb}mb.err
  done
fi
# DRAM traffic per launch with the caches left alone between kernels (ncu's default flushes them, which hides whether
# W really stays in L2 between K1 and K2h/K2): two metrics = one pass, no replay
if [ "$2" = traffic -o "$3" = traffic ]; then
  for wl in c2 c3 c4; do
    timeout 150 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      -k regex:wso_ -s 45 -c 36 --csv --log-file $OUT/traffic_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline > $OUT/traffic_$wl.log 2>&1
  done
fi
tail -n 3 $OUT/pytest_gpu.log; cat $OUT/lat_bench.jsonl $OUT/lat_bench_variants.txt 2>/dev/null; python tools/summ.py $OUT/bench_c?.json $OUT/var_*.json
