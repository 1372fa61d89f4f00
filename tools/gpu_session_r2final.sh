#!/bin/bash
# round-2 closing session on one B200: CEM + parity tests, the default bench line, then the ncu profiles
mkdir -p gpurun_out/fin
timeout 900 python -m pytest tests -m gpu -x -q -k "cem or CEM or parity_mode or train_kernels" > gpurun_out/fin/tests.log 2>&1
echo "tests rc=$?" > gpurun_out/fin/summary.txt
timeout 900 python bench.py > gpurun_out/fin/bench_1gpu.json 2> gpurun_out/fin/bench_1gpu.err
echo "bench rc=$?" >> gpurun_out/fin/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin/bench_ref.json 2> gpurun_out/fin/bench_ref.err
echo "ref rc=$?" >> gpurun_out/fin/summary.txt
bash tools/gpu_session_prof.sh > gpurun_out/fin/prof.log 2>&1
tail -5 gpurun_out/fin/tests.log
cat gpurun_out/fin/summary.txt gpurun_out/prof/summary.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/fin/bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('by_kernel'))
print({k: d.get(k) for k in ('forward', 'gan_step', 'zopt')})
PY
