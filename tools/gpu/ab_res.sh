#!/bin/bash
# debug build: pair kernel residual ring depth (SPK_PAIR_RES_SLOTS) on ResNet-50 and ResNet-18
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bench_parity.py -q -m gpu -x -k "not fp32" 2>&1 | tail -3 | cut -c1-300
for v in 2 4 2 4; do
  SPK_PAIR_RES_SLOTS=$v timeout 600 python bench.py --arch resnet50 --steps 30 --warmup 3 --no-cpu-baseline --e2e-bins 2 --profile-detail gpurun_out/pd_res_$v.tsv 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r50 res_slots=$v', round(d['value']), 'ms', round(d['ms_per_step'],4), 'dp', d['parity']['max_dp'], d['clocks']['reasons'])"
  grep "+res" gpurun_out/pd_res_$v.tsv | cut -c1-120
done
for v in 2 4; do
  SPK_PAIR_RES_SLOTS=$v timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --e2e-bins 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r18 res_slots=$v', round(d['value']), 'ms', round(d['ms_per_step'],4))"
done
