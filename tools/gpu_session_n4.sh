#!/bin/bash
# 4-GPU run of the driver's command line with the final library
mkdir -p gpurun_out/n4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4/b4.json 2> gpurun_out/n4/b4.err
echo "run N=4 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/n4/b4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('notes'), d.get('attempt'), {k: (d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('s_per_100_iters')) for k in ('forward', 'gan_step', 'zopt')})
PY
