#!/bin/bash
# Dev tool (under gpurun, one GPU): the evidence files of the round for the final tree - GPU tests, smoke, the default bench line
# (C2 with targets + cpu_baseline), C1 / C3 / C4, single-frame latency, reference arm, ncu launch list, --set full captures of C2 / C3
# summarised on the box (the reports are too big to bring both back), unflushed DRAM traffic.   usage: bash tools/gpu_r3z.sh TAG
TAG=${1:-r3z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/host.txt; free -g | head -2 >> $OUT/host.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
[ -x tools/lat_bench ] || g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include tools/lat_bench.cpp -o tools/lat_bench -L watersurfacerendering_b200 -lwsocean -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/watersurfacerendering_b200
for wl in c2 c1 c3 c4; do
  extra="--no-cpu-baseline"; [ $wl = c2 ] && extra=""
  timeout 600 python bench.py --workload $wl $extra > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref_c2.err; echo "reference arm rc=$?"
for n in 512 1024; do timeout 120 tools/lat_bench $n 2000 >> $OUT/lat_bench.jsonl 2>> $OUT/lat_bench.err; done
echo -n "WSO_K2_SPLIT=1 " >> $OUT/lat_variants.txt; WSO_K2_SPLIT=1 timeout 120 tools/lat_bench 512 2000 >> $OUT/lat_variants.txt 2>&1
echo -n "WSO_FRAME_GRAPH=0 " >> $OUT/lat_variants.txt; WSO_FRAME_GRAPH=0 timeout 120 tools/lat_bench 512 2000 >> $OUT/lat_variants.txt 2>&1
echo -n "WSO_FRAME_GRAPH=0 WSO_K2_SPLIT=1 " >> $OUT/lat_variants.txt; WSO_FRAME_GRAPH=0 WSO_K2_SPLIT=1 timeout 120 tools/lat_bench 512 2000 >> $OUT/lat_variants.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wso_ -s 45 -c 150 --csv --log-file $OUT/launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-targets > $OUT/launches_c2.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_lat512.csv tools/lat_bench 512 20 > /dev/null 2>&1
for wl in c2 c3; do
  skip=30; [ $wl = c3 ] && skip=12
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:wso_ -s $skip -c 3 -o $OUT/prof_$wl -f \
    python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-targets > $OUT/ncu_full_$wl.log 2>&1; echo "full $wl rc=$?"
  python tools/ncu_summary.py $OUT/prof_$wl.ncu-rep > $OUT/ncu_$wl.txt 2>&1
  for k in wso_pass1_kernel wso_heights_kernel wso_pass2_kernel; do
    echo "### $k" >> $OUT/ncu_mem_instr_$wl.txt; python tools/ncu_mem_instr.py $OUT/prof_$wl.ncu-rep $k >> $OUT/ncu_mem_instr_$wl.txt 2>&1
  done
done
rm -f $OUT/prof_c3.ncu-rep
bash tools/gpu_traffic.sh $TAG c2:d:64 c3:d:64 c4:d:64 c1:d:64 > $OUT/traffic_summary.txt 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/smoke.log; cat $OUT/lat_bench.jsonl $OUT/lat_variants.txt; python tools/summ.py $OUT/bench_c?.json; cut -c1-600 $OUT/bench_ref_c2.json; cat $OUT/traffic_summary.txt
