#!/bin/bash
mkdir -p gpurun_out/t
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t/tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/t/tests.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/t/bench.json 2> gpurun_out/t/bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/t/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('attempt'))
PY
