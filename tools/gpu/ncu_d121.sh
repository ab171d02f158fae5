#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --arch densenet121 --steps 2 --warmup 1 --no-cpu-baseline --e2e-bins 2"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 2 -o gpurun_out/r2_d121_conv_tc_pre -f $CMD > gpurun_out/ncu_d121.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_d121.log
