OUT=gpurun_out/r1f; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wso_ -s 30 -c 3 -o $OUT/prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_c2.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err
tail -3 $OUT/pytest_gpu.log; cut -c1-600 $OUT/bench_c2.json; ls -la $OUT
