#!/bin/bash
# Dev tool (under gpurun, one GPU): GPU test suite, then A/B of K2 packing from registers (in-tree) against the
# store-and-reload pack (variant library all_nopackregs) on every workload, and the single-tile latency.   usage: gpu_r3e.sh TAG
TAG=${1:-r3e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -n 3 $OUT/pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
V=$PWD/build/variants/libwsocean_all_notwpre.so
for m in 0 2; do
  WSO_WARP_CORE=$m timeout 200 python bench.py --workload c2 $B > $OUT/bench_c2_new_m$m.json 2> $OUT/bench_c2_new_m$m.err
  WSO_WARP_CORE=$m WSO_LIB_PATH=$V timeout 200 python bench.py --workload c2 $B > $OUT/bench_c2_old_m$m.json 2> $OUT/bench_c2_old_m$m.err
done
for wl in c3 c4 c1; do
  WSO_WARP_CORE=0 timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_new.json 2> $OUT/bench_${wl}_new.err
  WSO_WARP_CORE=0 WSO_LIB_PATH=$V timeout 200 python bench.py --workload $wl $B > $OUT/bench_${wl}_old.json 2> $OUT/bench_${wl}_old.err
done
python tools/summ.py $OUT/bench_*.json
[ -x tools/lat_bench ] || g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include tools/lat_bench.cpp -o tools/lat_bench -L watersurfacerendering_b200 -lwsocean -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/watersurfacerendering_b200
for n in 512 1024; do LD_LIBRARY_PATH=$PWD/watersurfacerendering_b200 timeout 120 tools/lat_bench $n 2000 2>&1 | tee -a $OUT/lat_bench.jsonl; done
