CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_hp_kernel -s 0 -c 1 -o gpurun_out/prof_n3 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
