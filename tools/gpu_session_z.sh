#!/bin/bash
mkdir -p gpurun_out/z
for v in "A=1" "ESR_ZERO_SCRATCH=1" "ESR_PDL=0" "ESR_POISON=1"; do
  echo "--- $v"
  env $v timeout 300 python tools/zopt_time.py 2>&1 | grep "^rep"
done 2>&1 | tee gpurun_out/z/variants.txt
