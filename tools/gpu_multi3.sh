#!/bin/bash
# Dev tool (under gpurun --gpus N): world-N slab parity tests + bench c5 in the direct exchange modes.  usage: gpu_multi3.sh TAG N
TAG=$1; N=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_slab.py -m gpu -x -q -k "multi_gpu and $N]" > $OUT/pytest_slab_multi.log 2>&1; echo "pytest rc=$?"; tail -n 3 $OUT/pytest_slab_multi.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
for fused in 1 3; do
  timeout 400 $RUN --workload c5 --steps 5 --warmup 3 --slab-fused $fused --no-cpu-baseline > $OUT/bench_c5_f${fused}_n$N.json 2> $OUT/bench_c5_f${fused}_n$N.err; echo "c5 fused=$fused rc=$?"
done
python - "$OUT" "$N" <<'PY'
import json, sys, glob
out, n = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(f"{out}/bench_*_n{n}.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d["roofline"]; nv = r.get("nvlink")
    print(f.split("/")[-1], "value", round(d["value"], 1), "us/tf", round(d["us_per_tile_frame"], 2), {k: round(v, 3) for k, v in r["phase_ms"].items()},
          "nvlink GB/s/dir", round(nv["achieved_gbs_per_dir"], 1) if nv else None)
PY
