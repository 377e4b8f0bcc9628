#!/bin/bash
# Dev tool (under gpurun --gpus P): 2-rank slab parity test + slab benches (c5s, c5; all-to-all and fused) + sharded c3/c4.
TAG=${1:-r1m}; P=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_slab.py -m gpu -x -q -k two_gpus > $OUT/pytest_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_2gpu.log
for wl in c5s c5; do for fused in 0 1; do
  timeout 900 $TR bench.py --gpus $P --workload $wl --slab-fused $fused --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${wl}_p${P}_f${fused}.json 2> $OUT/bench_${wl}_p${P}_f${fused}.err; echo "bench $wl fused=$fused rc=$?"
done; done
for wl in c2 c3 c4; do
  timeout 600 $TR bench.py --gpus $P --workload $wl --no-cpu-baseline > $OUT/bench_${wl}_p${P}.json 2> $OUT/bench_${wl}_p${P}.err; echo "bench $wl rc=$?"
done
tail -3 $OUT/pytest_2gpu.log; for f in $OUT/bench_c5*.json; do echo $f; cut -c1-700 $f; done
