#!/bin/bash
# parity at the benchmarked configurations + bench (with its own parity record) + the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bench_parity.py -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
grep -E "launch|fp32:|passed|failed|Error|assert" gpurun_out/t_parity.log | cut -c1-400 | tail -40
timeout 600 python bench.py --steps 20 --warmup 5 --profile-detail gpurun_out/prof_detail.tsv > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
tail -c 1500 gpurun_out/bench_ref.log
