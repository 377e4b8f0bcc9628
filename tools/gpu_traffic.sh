#!/bin/bash
# Dev tool (under gpurun, one GPU): DRAM bytes per kernel launch with the caches left alone between kernels
# (ncu --cache-control none), for a list of WL:MASK:MB[:EXTRAENV] configurations (MASK d = the default kernel choice).   usage: gpu_traffic.sh TAG cfg...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in "$@"; do
  IFS=: read wl m mb extra <<< "$cfg"
  name=${wl}_m${m}_mb${mb}${extra:+_$extra}
  maskenv="WSO_WARP_CORE=$m"; [ "$m" = d ] && maskenv="WSO_DEFAULT_KERNELS=1"   # d = the built-in choice per size
  env $extra $maskenv WSO_W_BUDGET_MB=$mb timeout 200 ncu --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum \
    -k regex:wso_ -s 45 -c 48 --csv --log-file $OUT/traffic_$name.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-targets > $OUT/traffic_$name.log 2>&1
  python - "$OUT/traffic_$name.csv" "$name" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
acc = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
for r in rows[1:]:
    k = r[ik].split("<")[0].replace("void ", "")
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if "byte" in u:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if r[im].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
        cnt[k] += 1
    acc[k][r[im]] += v
print(sys.argv[2])
for k in acc:
    n = cnt[k]
    a = acc[k]
    print(f"  {k:28s} launches {n:3d}  us/launch {a['gpu__time_duration.sum']/n:8.2f}  dram rd {a['dram__bytes_read.sum']/n/1e6:7.2f} MB  wr {a['dram__bytes_write.sum']/n/1e6:7.2f} MB  L2 rd {a['lts__t_sectors_op_read.sum']*32/n/1e6:7.1f} MB wr {a['lts__t_sectors_op_write.sum']*32/n/1e6:7.1f} MB")
PY
done
