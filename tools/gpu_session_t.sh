#!/bin/bash
mkdir -p gpurun_out/t
timeout 600 python -m pytest tests -m gpu -x -q -k "zopt or discriminator" > gpurun_out/t/tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/t/tests.log
