#!/bin/bash
# throughput against the batch size (ROIs per K2+K3 launch sequence); K1 always decodes 4096 ROIs per launch
for B in 256 512 1024; do
  timeout 300 python bench.py --batch $B --chunk-batches $((4096/B)) --steps $((25600/B)) --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > /tmp/b.json
  python -c "
import json
d=json.load(open('/tmp/b.json'))
print('batch', $B, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], round(d['roofline']['frac'],3))"
done
