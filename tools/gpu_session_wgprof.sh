#!/bin/bash
# ncu over the wgrad micro-benchmark: durations of every launch + full sections for three shapes
mkdir -p gpurun_out/wgp
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3_wgrad_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/wgp/w64 python tools/bench_wgrad.py > gpurun_out/wgp/w64.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3_wgrad_kernel --launch-skip 95 --launch-count 1 -o gpurun_out/wgp/w192 python tools/bench_wgrad.py > gpurun_out/wgp/w192.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/wgp/durations.csv python tools/bench_wgrad.py > gpurun_out/wgp/dur.log 2>&1
for r in w64 w192; do
  ncu -i gpurun_out/wgp/$r.ncu-rep --page raw --csv > gpurun_out/wgp/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/wgp/$r.ncu-rep --page source --csv > gpurun_out/wgp/$r.source.csv 2>/dev/null
  ncu -i gpurun_out/wgp/$r.ncu-rep --page details > gpurun_out/wgp/$r.details.txt 2>/dev/null
  rm -f gpurun_out/wgp/$r.ncu-rep
done
ls -la gpurun_out/wgp
