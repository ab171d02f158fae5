#!/bin/bash
# multi-GPU pass (arg1 = N GPUs, arg2 = archive bins, default 1000; 0 = skip the archive): BASELINE configs 2-5 under torchrun, one rank per GPU
N=${1:-8}
BINS=${2:-1000}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" == "1" ]; then TR="python"; fi
run() {  # name, command...
  local name=$1; shift
  timeout 600 "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name rc=$?"
  grep '^{' gpurun_out/$name.log | tail -1 > gpurun_out/$name.json
}
run r2_bench_r18_g$N $TR bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline
run r2_bench_r50_g$N $TR bench.py --gpus $N --arch resnet50 --steps 30 --warmup 3 --no-cpu-baseline --e2e-bins 24
run r2_bench_d121_g$N $TR bench.py --gpus $N --arch densenet121 --steps 30 --warmup 3 --no-cpu-baseline --e2e-bins 24
df -h /dev/shm | tail -1
[ "$BINS" != "0" ] && run r2_archive_g$N $TR tools/archive_bench.py --bins $BINS
python - <<PY
import json
for f in ['r2_bench_r18_g$N','r2_bench_r50_g$N','r2_bench_d121_g$N','r2_archive_g$N']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read())
        e=d.get('e2e') or {}
        print(f, round(d['value']), 'e2e', round(e.get('value',0)), e.get('stage_seconds_rank0'), d.get('first_open_to_last_csv_s'), d.get('stage_seconds_rank0'), 'ms', d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'))
    except Exception as ex:
        print(f, 'ERR', ex); print(open(f'gpurun_out/{f}.err').read()[-800:])
PY
