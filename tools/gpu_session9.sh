#!/bin/bash
OUT=gpurun_out/s10
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" > $OUT/summary.txt
tail -n 14 $OUT/pytest_gpu.log >> $OUT/summary.txt
timeout 300 ncu --kernel-name regex:"cem_" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/ncu_cem.csv python tools/stress_legs.py fwd --iters 1 --batch 16 --lr 256 --nb 1 > $OUT/ncu_cem.log 2>&1
echo "ncu rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
