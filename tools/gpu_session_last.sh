#!/bin/bash
mkdir -p gpurun_out/last
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke ok"
timeout 900 python bench.py > gpurun_out/last/bench_1gpu.json 2> gpurun_out/last/bench_1gpu.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/last/bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('by_kernel'))
print({k: (d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('s_per_100_iters')) for k in ('forward', 'gan_step', 'zopt')}, d.get('cpu_baseline', {}).get('value'), d.get('clocks'), d.get('attempt'))
PY
