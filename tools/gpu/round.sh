#!/bin/bash
# GPU pass: parity tests file by file (a CUDA fault poisons only its own process), smoke, bench, ncu
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_preprocess.py -q -m gpu > gpurun_out/t_pre.log 2>&1; echo "pre rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "stem" > gpurun_out/t_stem.log 2>&1; echo "stem rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "not stem" > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_network.py -q -m gpu > gpurun_out/t_net.log 2>&1; echo "net rc=$?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 100 --warmup 5 --profile-detail gpurun_out/prof_detail.tsv > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --arch resnet50 --steps 30 --warmup 3 --no-cpu-baseline --profile-detail gpurun_out/prof_detail_r50.tsv > gpurun_out/bench_r50.log 2>&1; echo "bench r50 rc=$?" >> gpurun_out/summary.txt
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?" >> gpurun_out/summary.txt
  timeout 600 $CMD > gpurun_out/ncu_plain2.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 19 -c 4 -o gpurun_out/prof_conv_tc -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
for f in t_pre t_stem t_conv t_net smoke bench bench_r50; do echo "=== $f"; tail -n 8 gpurun_out/$f.log | cut -c1-1500; done
