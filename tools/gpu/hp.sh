#!/bin/bash
# halo-pair kernel: TMEM bandwidth microbenchmark, conv parity tests, bench with in-step kernel table
mkdir -p gpurun_out
[ -x experiments/tmem_ld ] && timeout 120 ./experiments/tmem_ld
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_bench_parity.py -q -m gpu -s -k "bf16 or fp32_tc" 2>&1 | grep -E "launch|fp32_tc:|WARNING|passed|failed|Error|assert" | tail -20
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --profile-detail gpurun_out/prof_detail.tsv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'], d['parity']['max_dp'])
PY
cut -c1-150 gpurun_out/prof_detail.tsv
