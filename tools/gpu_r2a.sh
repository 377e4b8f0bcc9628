#!/bin/bash
# Dev tool (under gpurun, one GPU): parity tests, then A/B of the warp-per-line kernels against the CTA-per-line ones
# kernel by kernel (WSO_WARP_CORE bit mask: 1 = K1, 2 = K2h, 4 = K2).   usage: bash tools/gpu_r2a.sh TAG [masks...]
TAG=${1:-r2a}; shift
MASKS=${@:-"0 7 1 2 4"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -n 5 $OUT/pytest_gpu.log
for wl in c2 c3 c4; do
  for m in $MASKS; do
    WSO_WARP_CORE=$m timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${wl}_m$m.json 2> $OUT/bench_${wl}_m$m.err; echo "bench $wl mask $m rc=$?"
  done
done
python tools/summ.py $OUT/bench_*.json
