#!/bin/bash
# GPU pass: parity tests file by file (a CUDA fault poisons only its own process), smoke, bench (+ the reference arm)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_preprocess.py -q -m gpu > gpurun_out/t_pre.log 2>&1; echo "pre rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_network.py -q -m gpu > gpurun_out/t_net.log 2>&1; echo "net rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu_bench_parity.py -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 100 --warmup 5 --profile-detail gpurun_out/prof_detail.tsv > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
if [ "$1" == "all" ]; then
timeout 900 python bench.py --arch resnet50 --steps 30 --warmup 3 --no-cpu-baseline --profile-detail gpurun_out/prof_detail_r50.tsv > gpurun_out/bench_r50.log 2>&1; echo "bench r50 rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --arch densenet121 --steps 30 --warmup 3 --no-cpu-baseline --profile-detail gpurun_out/prof_detail_d121.tsv > gpurun_out/bench_d121.log 2>&1; echo "bench d121 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
for f in t_pre t_conv t_net t_parity smoke; do echo "=== $f"; grep -E "launch [0-9]+:|fp32:|passed|failed|Error|smoke" gpurun_out/$f.log | tail -n 12 | cut -c1-300; done
python - <<'PY'
import json
for f in ['bench','bench_r50','bench_d121','bench_ref']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],4), 'parity', d.get('parity'), 'roof', d.get('roofline',{}).get('frac'), {k: round(v,4) for k,v in d.get('kernel_ms_per_step',{}).items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
