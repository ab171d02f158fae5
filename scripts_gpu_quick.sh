#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_network.py tests/test_gpu_conv.py -q -m gpu -x > gpurun_out/t_net.log 2>&1; echo "net+conv rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 100 --warmup 5 --profile-detail gpurun_out/prof_detail.tsv > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --arch resnet50 --steps 30 --warmup 3 --no-cpu-baseline --profile-detail gpurun_out/prof_detail_r50.tsv > gpurun_out/bench_r50.log 2>&1; echo "bench r50 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 5 gpurun_out/t_net.log | cut -c1-600
python - <<'PY'
import json
for f in ['bench','bench_r50']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        print(f, round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms_per_step'], 'roof', round(d['roofline']['achieved']), d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/{f}.log').read()[-1500:])
PY
cut -c1-150 gpurun_out/prof_detail.tsv | head -5
