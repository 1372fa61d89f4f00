#!/bin/bash
OUT=gpurun_out/s8
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" > $OUT/summary.txt
tail -n 8 $OUT/pytest_gpu.log >> $OUT/summary.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/summary.txt
timeout 300 ncu --set full --kernel-name regex:cem_up_add -c 1 --clock-control none -o $OUT/up_add python tools/stress_legs.py fwd --iters 1 --batch 16 --lr 256 --nb 1 > $OUT/ncu_up.log 2>&1
ncu -i $OUT/up_add.ncu-rep --page details --csv > $OUT/up_add_details.csv 2>/dev/null
echo "ncu rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
