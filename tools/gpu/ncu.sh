#!/bin/bash
# ncu pass: launch list of one bench run + full capture of selected kernels (arg1 = kernel regex, arg2 = skip, arg3 = count)
mkdir -p gpurun_out
KREGEX=${1:-conv3x3_halo}
SKIP=${2:-8}
COUNT=${3:-4}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 $CMD > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT -o gpurun_out/prof_$KREGEX -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
