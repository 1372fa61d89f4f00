#!/bin/bash
mkdir -p gpurun_out/t
timeout 900 python -m pytest tests -m gpu -x -q -k "pixelshuffle or zopt or pixel_shuffle" > gpurun_out/t/tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/t/tests.log
