#!/bin/bash
# all parity tests, then the three benches (short)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -3 | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_parity.py -q -m gpu -s 2>&1 | grep -E "launch 256:|fp32|passed|failed|Error" | tail -12 | cut -c1-200
for a in resnet18 resnet50 densenet121; do
timeout 600 python bench.py --arch $a --steps 40 --warmup 3 --no-cpu-baseline --e2e-bins 4 --profile-detail gpurun_out/pd_$a.tsv 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a', round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'dp', d['parity']['max_dp'], d['clocks']['reasons'])"
done
timeout 600 python bench.py --precision fp32_tc --steps 20 --warmup 3 --no-cpu-baseline --e2e-bins 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r18 fp32_tc', round(d['value']), 'ms', round(d['ms_per_step'],4), 'dp', d['parity']['max_dp'])"
