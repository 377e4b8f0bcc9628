#!/bin/bash
# Dev tool (under gpurun, one GPU): GPU test suite, then the single-frame path with and without the frame graph:
# tools/lat_bench (C ABI) at 512 / 1024 and bench.py --workload c1 (Python binding).   usage: gpu_r3l.sh TAG
TAG=${1:-r3l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -n 5 $OUT/pytest_gpu.log
[ -x tools/lat_bench ] || g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include tools/lat_bench.cpp -o tools/lat_bench -L watersurfacerendering_b200 -lwsocean -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/watersurfacerendering_b200
for g in 1 0; do
  for n in 512 1024 2048; do
    echo -n "graph=$g " >> $OUT/lat_bench.txt
    WSO_FRAME_GRAPH=$g timeout 120 tools/lat_bench $n 2000 >> $OUT/lat_bench.txt 2>> $OUT/lat_bench.err
  done
  WSO_FRAME_GRAPH=$g timeout 200 python bench.py --workload c1 --no-cpu-baseline --no-targets > $OUT/bench_c1_graph$g.json 2> $OUT/bench_c1_graph$g.err
done
cat $OUT/lat_bench.txt
python - <<P
import json
for g in (1,0):
    d=json.load(open("$OUT/bench_c1_graph%d.json"%g)); print("bench c1 graph",g, round(d["us_per_tile_frame"],2),"us/step; e2e",round(d["e2e"]["value"]))
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_lat512.csv tools/lat_bench 512 20 > /dev/null 2>&1; tail -8 $OUT/launches_lat512.csv
