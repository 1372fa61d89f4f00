#!/bin/bash
# smoke() + full GPU suite + the default bench line + the reference arm
mkdir -p gpurun_out/full
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full/smoke.log 2>&1
echo "smoke rc=$?" > gpurun_out/full/summary.txt
grep "smoke ok" gpurun_out/full/smoke.log >> gpurun_out/full/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full/tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/full/summary.txt
tail -3 gpurun_out/full/tests.log >> gpurun_out/full/summary.txt
timeout 900 python bench.py > gpurun_out/full/bench_1gpu.json 2> gpurun_out/full/bench_1gpu.err
echo "bench rc=$?" >> gpurun_out/full/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/full/bench_ref.json 2> gpurun_out/full/bench_ref.err
echo "ref rc=$?" >> gpurun_out/full/summary.txt
python - <<'PY' >> gpurun_out/full/summary.txt
import json
d = json.loads(open('gpurun_out/full/bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('by_kernel'))
print({k: (d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('s_per_100_iters')) for k in ('forward', 'gan_step', 'zopt')}, d.get('cpu_baseline'), d.get('clocks'))
PY
cat gpurun_out/full/summary.txt
