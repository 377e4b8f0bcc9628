#!/bin/bash
# Dev tool (under gpurun --gpus 2): 2-rank slab parity tests, the 2048^2 slab bench in both exchange modes, sharded c2.
TAG=${1:-r1q}; P=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29511"
timeout 400 python -m pytest tests/test_slab.py tests/test_sharding.py -m gpu -x -q > $OUT/pytest_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_2gpu.log
for fused in 0 1; do
  timeout 200 $TR bench.py --gpus $P --workload c5s --slab-fused $fused --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_c5s_p${P}_f${fused}.json 2> $OUT/bench_c5s_p${P}_f${fused}.err; echo "bench c5s fused=$fused rc=$?"
done
timeout 300 $TR bench.py --gpus $P --workload c2 --no-cpu-baseline > $OUT/bench_c2_p${P}.json 2> $OUT/bench_c2_p${P}.err; echo "bench c2 rc=$?"
tail -n 3 $OUT/pytest_2gpu.log; for f in $OUT/bench_*.json; do echo $f; cut -c1-400 $f; done
