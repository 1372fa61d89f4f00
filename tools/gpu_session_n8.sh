#!/bin/bash
# three consecutive 8-GPU runs of the driver's command line, then one 2-GPU run
OUT=gpurun_out/n8
mkdir -p $OUT
: > $OUT/summary.txt
for i in 1 2 3; do
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/b8_$i.json 2> $OUT/b8_$i.err
  echo "run $i rc=$?" >> $OUT/summary.txt
  python - <<PY >> $OUT/summary.txt 2>&1
import json
try:
    d = json.loads(open('$OUT/b8_$i.json').read().strip().splitlines()[-1])
    print('  value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', d.get('forward', {}).get('ms_per_step'), 'gan', d.get('gan_step', {}).get('ms_per_step'), 'zopt', d.get('zopt', {}).get('s_per_100_iters'), 'notes', d.get('notes'))
except Exception as e:
    print('  no line:', e)
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/b2.json 2> $OUT/b2.err
echo "run N=2 rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
