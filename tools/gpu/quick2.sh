#!/bin/bash
# conv + network parity, then the bench with the in-step kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_parity.py -q -m gpu -s 2>&1 | grep -E "launch [0-9]+:|fp32|passed|failed|Error" | tail -14 | cut -c1-200
timeout 900 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-bins 8 --profile-detail gpurun_out/prof_detail.tsv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3), 'dominant', round(r['dominant']['frac'],3), r['dominant']['ms_per_launch'], d['kernel_ms_per_step'], d['parity']['max_dp'], d['clocks'])
PY
cut -c1-150 gpurun_out/prof_detail.tsv
