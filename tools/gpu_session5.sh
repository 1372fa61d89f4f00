#!/bin/bash
OUT=gpurun_out/s6
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" > $OUT/summary.txt
tail -n 15 $OUT/pytest_gpu.log >> $OUT/summary.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
