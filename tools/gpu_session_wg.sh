#!/bin/bash
mkdir -p gpurun_out/wg
timeout 300 python tools/check_wgrad.py > gpurun_out/wg/debug.txt 2>&1
echo "debug rc=$?" > gpurun_out/wg/summary.txt
cat gpurun_out/wg/debug.txt >> gpurun_out/wg/summary.txt
if grep -q "nan [1-9]\|Error\|error" gpurun_out/wg/debug.txt; then cat gpurun_out/wg/summary.txt; exit 0; fi
timeout 600 python -m pytest tests -m gpu -x -q -k "wgrad or train_gpu or parity_mode or determin or pixelshuffle" > gpurun_out/wg/tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/wg/summary.txt
tail -3 gpurun_out/wg/tests.log >> gpurun_out/wg/summary.txt
timeout 300 python tools/bench_wgrad.py > gpurun_out/wg/v2.txt 2>&1
cat gpurun_out/wg/v2.txt >> gpurun_out/wg/summary.txt
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/wg/bench.json 2> gpurun_out/wg/bench.err
echo "bench rc=$?" >> gpurun_out/wg/summary.txt
python - <<'PY' >> gpurun_out/wg/summary.txt
import json
d = json.loads(open('gpurun_out/wg/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('by_kernel'))
PY
cat gpurun_out/wg/summary.txt
