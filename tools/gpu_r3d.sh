#!/bin/bash
# Dev tool (under gpurun, one GPU): full GPU test suite, persistent K1/K2 at 2048^2 (1 CTA per SM in K1) with two chunk sizes,
# the record-prefetch variant, slab K1 launch order at 16384^2 on one device.   usage: gpu_r3d.sh TAG
TAG=${1:-r3d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -n 3 $OUT/pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for mb in 64 140; do for m in 0 16 64 80; do
  WSO_WARP_CORE=$m WSO_W_BUDGET_MB=$mb timeout 200 python bench.py --workload c3 $B > $OUT/bench_c3_m${m}_mb$mb.json 2> $OUT/bench_c3_m${m}_mb$mb.err
done; done
WSO_WARP_CORE=0 WSO_LIB_PATH=$PWD/build/variants/libwsocean_n11_pf11.so timeout 200 python bench.py --workload c3 $B > $OUT/bench_c3_pf11.json 2> $OUT/bench_c3_pf11.err
python tools/summ.py $OUT/bench_c3*.json
timeout 300 python bench.py --workload c5 --steps 3 --warmup 2 --no-cpu-baseline > $OUT/c5_w1_new.json 2> $OUT/c5_w1_new.err
WSO_LIB_PATH=$PWD/build/variants/libwsocean_slab_xmajor.so timeout 300 python bench.py --workload c5 --steps 3 --warmup 2 --no-cpu-baseline > $OUT/c5_w1_xmajor.json 2> $OUT/c5_w1_xmajor.err
for f in $OUT/c5_w1_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d.get('ms_per_step'), d.get('phase_ms') or d.get('config',{}).get('phase_ms') or {k:v for k,v in d.items() if 'phase' in k})
"; done; tail -n 3 $OUT/c5_w1_*.err
