#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_parity.py -q -m gpu -k "densenet or d121" -s 2>&1 | grep -E "launch [0-9]+:|fp32|passed|failed|Error" | tail -8 | cut -c1-200
timeout 600 python bench.py --arch densenet121 --steps 30 --warmup 3 --no-cpu-baseline --e2e-bins 4 --profile-detail gpurun_out/pd_d121.tsv 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('d121', round(d['value']), 'ms', round(d['ms_per_step'],4), 'dp', d['parity']['max_dp'], d['clocks']['reasons'], d['kernel_ms_per_step'])"
grep "56x56 out 56x56" gpurun_out/pd_d121.tsv | cut -c1-130 | head -8
