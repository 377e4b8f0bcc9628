#!/bin/bash
# Dev tool (under gpurun --gpus N): multi-rank slab parity tests, then bench c5 (16384^2) in both exchange modes and the
# sharded workloads.   usage: bash tools/gpu_multi2.sh TAG N [skiptests]
TAG=$1; N=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$3" != skiptests ]; then
  timeout 900 python -m pytest tests/test_slab.py -m gpu -x -q -k "multi_gpu and ${3:-vs_oracle}" > $OUT/pytest_slab_multi.log 2>&1; echo "pytest rc=$?"; tail -n 4 $OUT/pytest_slab_multi.log
fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
for fused in 0 1; do
  timeout 600 $RUN --workload c5 --steps 5 --warmup 3 --slab-fused $fused --no-cpu-baseline > $OUT/bench_c5_f${fused}_n$N.json 2> $OUT/bench_c5_f${fused}_n$N.err; echo "c5 fused=$fused rc=$?"
done
for wl in c4 c2; do
  timeout 400 $RUN --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${wl}_n$N.json 2> $OUT/bench_${wl}_n$N.err; echo "$wl rc=$?"
done
python - "$OUT" "$N" <<'PY'
import json, sys, glob
out, n = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(f"{out}/bench_*_n{n}.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d["roofline"]
    extra = {k: round(v, 3) for k, v in r.get("phase_ms", {}).items()} if "phase_ms" in r else {k: round(v, 3) for k, v in r["kernel_ms"].items()}
    nv = r.get("nvlink")
    print(f.split("/")[-1], "value", round(d["value"], 1), "us/tf", round(d["us_per_tile_frame"], 2), "e2e", round(d["e2e"]["value"], 1), extra,
          "nvlink GB/s/dir", round(nv["achieved_gbs_per_dir"], 1) if nv else None, "e2e probe", (d["e2e"].get("d2h_probe") or {}).get("gbs_per_gpu"))
PY
