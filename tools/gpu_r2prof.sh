#!/bin/bash
# Dev tool (under gpurun, one GPU): ncu --set full capture of one launch of each hot-path kernel.
# usage: gpu_r2prof.sh TAG WL [MASK]
TAG=${1:-r2p}; WL=${2:-c2}; MASK=${3:-7}
OUT=gpurun_out/$TAG; mkdir -p $OUT
WSO_WARP_CORE=$MASK timeout 900 ncu --set full --clock-control none --import-source on -k regex:wso_ -s 30 -c 3 -o $OUT/prof_$WL -f python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$WL.log 2>&1
echo "ncu rc=$?"; ls -la $OUT
