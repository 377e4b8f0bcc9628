#!/bin/bash
# Dev tool (under gpurun): bench every build/variants/libwsocean_n<LOGN>_*.so on one workload.
#   bash tools/sweep_run.sh LOGN WORKLOAD TAG
LOGN=$1; WL=$2; TAG=${3:-sweep}
mkdir -p gpurun_out/$TAG
for so in build/variants/libwsocean_n${LOGN}_*.so; do
  name=$(basename $so .so); name=${name#libwsocean_}
  WSO_LIB_PATH=$PWD/$so timeout 120 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/$TAG/$name.json 2> gpurun_out/$TAG/$name.err
done
python tools/summ.py gpurun_out/$TAG/*.json
