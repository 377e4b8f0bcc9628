#!/bin/bash
# Dev tool (under gpurun, one GPU): the evidence files of the round for the DEFAULT configuration -
#   launch list of the bench command (per-launch device times: shares, not absolutes), --set full capture of the three kernels,
#   unflushed DRAM traffic per workload.     usage: bash tools/gpu_final_profiles.sh TAG
TAG=${1:-r2z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wso_ -s 45 -c 150 --csv --log-file $OUT/launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-targets > $OUT/launches_c2.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wso_ -s 30 -c 3 -o $OUT/prof_c2 -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-targets > $OUT/ncu_full_c2.log 2>&1; echo "full c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wso_ -s 12 -c 3 -o $OUT/prof_c3 -f \
  python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-targets > $OUT/ncu_full_c3.log 2>&1; echo "full c3 rc=$?"
bash tools/gpu_traffic.sh $TAG c2:d:64 c3:d:64 c4:d:64 c1:d:64
