#!/bin/bash
# Dev tool (under gpurun, one GPU): parity tests, A/B of the two kernel sets (WSO_WARP_CORE mask), chunk budget and the
# variant libraries in build/variants/ (name n<LOGN>_*), all on the warp-per-line kernels.   usage: gpu_r2b.sh TAG
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
WSO_WARP_CORE=7 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -n 3 $OUT/pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline"
for m in 7 1 2 4; do WSO_WARP_CORE=$m timeout 200 python bench.py --workload c2 $B > $OUT/bench_c2_m$m.json 2> $OUT/bench_c2_m$m.err; done
for wl in c3 c4; do for m in 7; do WSO_WARP_CORE=$m timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_m$m.json 2> $OUT/bench_${wl}_m$m.err; done; done
for mb in 100; do WSO_WARP_CORE=7 WSO_W_BUDGET_MB=$mb timeout 200 python bench.py --workload c2 $B > $OUT/bench_c2_m7_mb$mb.json 2> $OUT/bench_c2_m7_mb$mb.err; done
for so in build/variants/libwsocean_n*.so; do
  [ -f $so ] || continue
  name=$(basename $so .so); name=${name#libwsocean_}
  case $name in n9_*) wl=c4;; n10_*) wl=c2;; n11_*) wl=c3;; *) wl=c2;; esac
  WSO_WARP_CORE=7 WSO_LIB_PATH=$PWD/$so timeout 100 python bench.py --workload $wl $B > $OUT/var_${name}.json 2> $OUT/var_${name}.err
  WSO_WARP_CORE=7 WSO_W_BUDGET_MB=50 WSO_LIB_PATH=$PWD/$so timeout 100 python bench.py --workload $wl $B > $OUT/var_${name}_mb50.json 2> $OUT/var_${name}_mb50.err
done
python tools/summ.py $OUT/bench_*.json $OUT/var_*.json
