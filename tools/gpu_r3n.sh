#!/bin/bash
# Dev tool (under gpurun --gpus N): C5 (direct peer stores), sharded C2 and C4 at N GPUs with the final tree.   usage: gpu_r3n.sh TAG N
TAG=$1; N=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
timeout 200 $RUN --workload c5 --steps 5 --warmup 3 --slab-fused 1 --no-cpu-baseline > $OUT/bench_c5_f1_n$N.json 2> $OUT/bench_c5_f1_n$N.err; echo "c5 rc=$?"
timeout 120 $RUN --workload c2 --no-cpu-baseline --no-targets > $OUT/bench_c2_n$N.json 2> $OUT/bench_c2_n$N.err; echo "c2 rc=$?"
timeout 120 $RUN --workload c4 --no-cpu-baseline --no-targets > $OUT/bench_c4_n$N.json 2> $OUT/bench_c4_n$N.err; echo "c4 rc=$?"
timeout 120 $RUN --workload c3 --no-cpu-baseline --no-targets > $OUT/bench_c3_n$N.json 2> $OUT/bench_c3_n$N.err; echo "c3 rc=$?"
for f in $OUT/bench_*_n$N.json; do echo $f; cut -c1-330 $f; done
python - "$OUT" "$N" <<'PY'
import json, sys
out, n = sys.argv[1], sys.argv[2]
d = json.loads([l for l in open(f"{out}/bench_c5_f1_n{n}.json") if l.startswith("{")][-1])
r = d["roofline"]; nv = r.get("nvlink")
print("c5", round(d["us_per_tile_frame"], 1), {k: round(v, 3) for k, v in r["phase_ms"].items()}, "nvlink GB/s/dir", round(nv["achieved_gbs_per_dir"], 1) if nv else None)
PY
