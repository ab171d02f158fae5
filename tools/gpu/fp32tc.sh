#!/bin/bash
# FP32_TC (fp32-level accuracy on the 16-bit tensor cores, the CLI default): parity on every backbone, overflow test, throughput.
# Debug build: SPK_SPLIT_CHUNK=<k blocks per TMEM accumulation chunk> sweeps the accumulation chunk.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network.py -q -m gpu 2>&1 | tail -6 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_bench_parity.py -q -m gpu -s -k "fp32_tc" 2>&1 | grep -E "fp32_tc:|passed|failed|Error|assert" | tail -12 | cut -c1-300
for a in resnet18 resnet50 densenet121; do
timeout 600 python bench.py --arch $a --precision fp32_tc --steps 20 --warmup 3 --no-cpu-baseline --e2e-bins 16 > gpurun_out/r2_bench_fp32tc_$a.json 2> gpurun_out/bench_fp32tc_$a.err; echo "bench $a rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_fp32tc_$a.json').read().strip().splitlines()[-1])
print('$a', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],4), d['kernel_ms_per_step'], d['parity'])
PY
done
