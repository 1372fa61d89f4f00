#!/bin/bash
mkdir -p gpurun_out/cvp
timeout 300 python tools/bench_conv.py > gpurun_out/cvp/times.txt 2>&1
cat gpurun_out/cvp/times.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3_rows_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/cvp/c64 python tools/bench_conv.py > gpurun_out/cvp/c64.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3_rows_kernel --launch-skip 72 --launch-count 1 -o gpurun_out/cvp/c160 python tools/bench_conv.py > gpurun_out/cvp/c160.log 2>&1
for r in c64 c160; do
  ncu -i gpurun_out/cvp/$r.ncu-rep --page source --csv > gpurun_out/cvp/$r.source.csv 2>/dev/null
  ncu -i gpurun_out/cvp/$r.ncu-rep --page details > gpurun_out/cvp/$r.details.txt 2>/dev/null
  rm -f gpurun_out/cvp/$r.ncu-rep
done
ls -la gpurun_out/cvp
