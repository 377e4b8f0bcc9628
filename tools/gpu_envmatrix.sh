#!/bin/bash
# Dev tool (under gpurun, one GPU): bench.py under a list of environment settings.  usage: gpu_envmatrix.sh TAG WL "ENV1=.. ENV2=.." ...
TAG=$1; WL=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 200 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-targets > $OUT/bench_${WL}_$i.json 2> $OUT/bench_${WL}_$i.err
  python - "$OUT/bench_${WL}_$i.json" "$cfg" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r = d["roofline"]
    print(f"{sys.argv[2]:60s} {d['us_per_tile_frame']:8.2f} us/tf  {d['value']:10.0f} tf/s  kernels {({k: round(v*1e3/(d['steps']*d['config']['frames_per_step_per_gpu']),2) for k,v in r['kernel_ms'].items()})}")
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
done
