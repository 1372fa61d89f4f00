#!/bin/bash
mkdir -p gpurun_out/t
timeout 900 python -m pytest tests -m gpu -x -q -k "train_gpu or discriminator or wgan or zopt or checkpoint" > gpurun_out/t/tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/t/tests.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/t/bench.json 2> gpurun_out/t/bench.err
echo "bench rc=$?"
ESR_SYNC_INPUT=1 ESR_SYNC_LOGS=1 timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/t/bench_sync.json 2> gpurun_out/t/bench_sync.err
python - <<'PY'
import json
for f in ('bench', 'bench_sync'):
    d = json.loads(open('gpurun_out/t/%s.json' % f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
