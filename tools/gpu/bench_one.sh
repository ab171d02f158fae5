mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 5 --profile-detail gpurun_out/prof_detail.tsv --profile-events gpurun_out/prof_events.tsv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['stage_seconds_rank0'], d['e2e']['seconds'], 'ms', d['ms_per_step'])
print(d['roofline'])
print(d['roofline_preprocess'])
print(d['kernel_ms_per_step'], d['parity'])
PY
cat gpurun_out/prof_detail.tsv | cut -c1-170
