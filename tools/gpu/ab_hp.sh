#!/bin/bash
# debug build: A/B of the halo-pair kernel's epilogue traffic (SPK_HP_DBG), per-launch event table
for v in 0 1 2 3; do
  SPK_HP_DBG=$v timeout 300 python bench.py --steps 32 --warmup 5 --no-cpu-baseline --profile-detail gpurun_out/prof_hp_$v.tsv > gpurun_out/bench_hp_$v.log 2>&1
  echo "== SPK_HP_DBG=$v"; grep "halo pair" gpurun_out/prof_hp_$v.tsv | cut -c1-130
done
