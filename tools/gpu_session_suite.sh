#!/bin/bash
# the GPU test suite alone
mkdir -p gpurun_out/suite
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/suite/tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/suite/tests.log
