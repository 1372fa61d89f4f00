#!/bin/bash
OUT=gpurun_out/s2
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" > $OUT/summary.txt
tail -n 30 $OUT/pytest_gpu.log >> $OUT/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/summary.txt
ESR_PRECISION=parity timeout 300 python tools/stress_legs.py train --iters 6 --batch 16 --lr 256 > $OUT/parity_train.log 2>&1
echo "parity train rc=$?" >> $OUT/summary.txt
tail -n 2 $OUT/parity_train.log >> $OUT/summary.txt
ESR_PRECISION=parity timeout 300 python tools/stress_legs.py fwd --iters 10 --batch 16 --lr 256 > $OUT/parity_fwd.log 2>&1
tail -n 1 $OUT/parity_fwd.log >> $OUT/summary.txt
cat $OUT/summary.txt
