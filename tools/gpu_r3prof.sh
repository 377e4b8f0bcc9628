#!/bin/bash
# Dev tool (under gpurun, one GPU): ncu --set full capture (with source) of the default kernels.  usage: gpu_r3prof.sh TAG WORKLOAD SKIP COUNT [KERNEL_REGEX]
# (one workload per call: gpurun brings back at most 64 MiB)
TAG=${1:-r3p}; WL=${2:-c2}; SKIP=${3:-30}; CNT=${4:-3}; KR=${5:-wso_}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c $CNT -o $OUT/prof_$WL -f \
  python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-targets > $OUT/ncu_full_$WL.log 2>&1; echo "full $WL rc=$?"
ls -la $OUT
