#!/bin/bash
# Dev tool (under gpurun, one GPU): parity tests + smoke + bench of every single-GPU workload + A/B of the variant
# libraries in build/variants/ + single-tile latency + ncu launch list + one ncu --set full capture.
# usage: bash tools/gpu_s8.sh TAG [noncu]
TAG=${1:-r1h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
[ -x tools/lat_bench ] || make -C examples lat_bench > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/host.txt
timeout 420 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
for wl in c2 c3 c4 c1; do
  extra="--no-cpu-baseline"; [ $wl = c2 ] && extra=""
  timeout 200 python bench.py --workload $wl $extra > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"
done
for so in build/variants/libwsocean_*.so; do
  [ -f $so ] || continue
  name=$(basename $so .so); name=${name#libwsocean_}
  for wl in c2 c3; do
    WSO_LIB_PATH=$PWD/$so timeout 100 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/var_${name}_$wl.json 2> $OUT/var_${name}_$wl.err
  done
done
for n in 512 1024; do timeout 60 tools/lat_bench $n 2000 >> $OUT/lat_bench.jsonl 2>> $OUT/lat_bench.err; done
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref_c2.err
if [ "$2" != noncu ]; then
  timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
  timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file $OUT/launches_lat512.csv tools/lat_bench 512 20 > /dev/null 2>&1
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:wso_ -s 30 -c 3 -o $OUT/prof_c2 \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_c2.log 2>&1
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:wso_ -s 12 -c 3 -o $OUT/prof_c3 \
     python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_c3.log 2>&1
fi
tail -n 3 $OUT/pytest_gpu.log; cat $OUT/lat_bench.jsonl; python tools/summ.py $OUT/bench_c?.json $OUT/var_*.json
