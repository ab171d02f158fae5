#!/bin/bash
# A/B at STEP level, alternating in one call on one box: the transposed stem (default) against the old two-pass kernel
mkdir -p gpurun_out
for i in 1 2 3; do
  for v in t hilo; do
    SPK_STEM=$v timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],4), d['clocks'])"
  done
done
