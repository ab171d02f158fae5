#!/bin/bash
# round-2 evidence pass at HEAD: ncu launch list of one bench run, then full captures of the kernels the roofline lines name
# (each only after the plain command has exited 0).  Outputs gpurun_out/r2_*.ncu-rep + launches; summarised by
# tools/ncu_to_profiles.py into profiles/r2_*.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-bins 4"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/r2_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
# launches in step order: the first two conv3x3_hp launches are layer1.0 conv1 (no residual) and conv2 (+residual)
cap conv3x3_hp64 'conv3x3_hp_kernel' 0 2
cap stem_pool_p 'stem_pool_p_kernel' 0 1
[ "$1" == "short" ] && { ls -la gpurun_out/r2_*; exit 0; }
cap conv_pair 'conv_pair_kernel' 0 6
cap preprocess_u8 'preprocess_u8_kernel' 0 1
cap head 'head_kernel' 0 1
ls -la gpurun_out/r2_*
