#!/bin/bash
# round-2 profiles: launch list of the headline step, `--set full` of the tensor-core kernels in a steady-state training step
OUT=gpurun_out/prof
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6500 --csv --log-file $OUT/launches_train.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $OUT/launches_train.log 2>&1
echo "launch list rc=$?" > $OUT/summary.txt
K='regex:conv3x3_rows_kernel|conv3x3_wgrad_kernel'
timeout 600 ncu --set full --clock-control none -k "$K" --launch-skip 2130 --launch-count 12 -o $OUT/fwd_full python tools/stress_legs.py train --iters 3 --batch 16 --lr 256 > $OUT/fwd_full.log 2>&1
echo "fwd full rc=$?" >> $OUT/summary.txt
timeout 600 ncu --set full --clock-control none -k "$K" --launch-skip 2500 --launch-count 24 -o $OUT/bwd_full python tools/stress_legs.py train --iters 3 --batch 16 --lr 256 > $OUT/bwd_full.log 2>&1
echo "bwd full rc=$?" >> $OUT/summary.txt
timeout 300 ncu --set full --clock-control none -k 'regex:cem_' -c 3 -o $OUT/cem_full python tools/stress_legs.py fwd --iters 1 --batch 16 --lr 256 --nb 1 > $OUT/cem_full.log 2>&1
for r in fwd_full bwd_full cem_full; do
  ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/$r.raw.csv 2>/dev/null
  rm -f $OUT/$r.ncu-rep
done
ls -la $OUT >> $OUT/summary.txt
cat $OUT/summary.txt
