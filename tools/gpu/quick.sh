#!/bin/bash
# quick GPU pass: conv + network parity, then a bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_network.py -q -m gpu > gpurun_out/t_net.log 2>&1; echo "net rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --profile-detail gpurun_out/prof_detail.tsv > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --arch resnet50 --steps 30 --warmup 3 --no-cpu-baseline --profile-detail gpurun_out/prof_detail_r50.tsv > gpurun_out/bench_r50.log 2>&1; echo "bench r50 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 6 gpurun_out/t_conv.log | cut -c1-400
tail -n 6 gpurun_out/t_net.log | cut -c1-400
python - <<'PY'
import json
for f in ['bench','bench_r50']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        print(f, round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v,4) for k,v in d['kernel_ms_per_step'].items()}, 'roof', round(d['roofline']['achieved']), round(d['roofline']['frac'],3))
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/{f}.log').read()[-1500:])
PY
cut -c1-150 gpurun_out/prof_detail.tsv
