#!/bin/bash
# debug build: the persistent stem (stem_pool_p_kernel) against the CTA-per-strip kernel (SPK_STEM_STRIPS=1): parity, then steps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x -k "stem" 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_parity.py -q -m gpu -x -k "bf16" 2>&1 | tail -4 | cut -c1-300
for v in p strips p; do
  if [ $v == strips ]; then export SPK_STEM_STRIPS=1; else unset SPK_STEM_STRIPS; fi
  timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-bins 4 --profile-detail gpurun_out/pd_stem_$v.tsv > gpurun_out/bench_stem_$v.log 2> gpurun_out/bench_stem_$v.err; echo "bench $v rc=$?"
  tail -c 300 gpurun_out/bench_stem_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_stem_$v.log').read().strip().splitlines()[-1])
print('$v', round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['kernel_ms_per_step'], d['parity']['max_dp'], d['clocks'])
PY
done
SPK_STEM_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-bins 2 2>&1 | grep "stem_p" | head -4
