#!/bin/bash
# Dev tool (under gpurun, one GPU): parity of every kernel-choice mask, then A/B of the persistent forms of K1 / K2
# (WSO_WARP_CORE bits 4 / 6) on C2 / C3 / C4.   usage: gpu_r2p.sh TAG
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -n 3 $OUT/pytest_gpu.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-targets"
for m in 2 18 66 82; do WSO_WARP_CORE=$m timeout 200 python bench.py --workload c2 $B > $OUT/bench_c2_m$m.json 2> $OUT/bench_c2_m$m.err; done
for m in 0 64; do WSO_WARP_CORE=$m timeout 200 python bench.py --workload c3 $B > $OUT/bench_c3_m$m.json 2> $OUT/bench_c3_m$m.err; done
for m in 0 16 64 80; do WSO_WARP_CORE=$m timeout 200 python bench.py --workload c4 $B > $OUT/bench_c4_m$m.json 2> $OUT/bench_c4_m$m.err; done
python tools/summ.py $OUT/bench_*.json
