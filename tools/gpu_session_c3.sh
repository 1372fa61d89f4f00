#!/bin/bash
mkdir -p gpurun_out/c3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3/launches_c3.csv python tools/profile_c3.py > gpurun_out/c3/c3.log 2>&1
echo "c3 list rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c3/bench.json 2> gpurun_out/c3/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c3/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
print({k: (d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('s_per_100_iters')) for k in ('forward', 'gan_step', 'zopt')})
PY
