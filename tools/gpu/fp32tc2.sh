#!/bin/bash
# debug build: FP32_TC with the fp16 hi + lo SplitF format: parity everywhere, error and speed against the accumulation chunk
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network.py -q -m gpu -x 2>&1 | tail -6 | cut -c1-300
for c in 2 8 1000; do
  echo "== SPK_SPLIT_CHUNK=$c"
  SPK_SPLIT_CHUNK=$c timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_network.py -q -m gpu -s -k "fp32_tc" 2>&1 | grep -E "fp32_tc:|passed|failed" | tail -8 | cut -c1-200
  SPK_SPLIT_CHUNK=$c timeout 600 python bench.py --precision fp32_tc --steps 20 --warmup 3 --no-cpu-baseline --e2e-bins 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r18', round(d['value']), d['parity']['max_dp'])"
done
