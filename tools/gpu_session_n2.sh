#!/bin/bash
# 2-GPU run of the driver's command line with the final library
mkdir -p gpurun_out/n2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2/b2.json 2> gpurun_out/n2/b2.err
echo "run N=2 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/n2/b2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('notes'), {k: (d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('s_per_100_iters')) for k in ('forward', 'gan_step', 'zopt')})
PY
