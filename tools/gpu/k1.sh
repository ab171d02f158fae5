mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py -q -m gpu -x > gpurun_out/t_pre.log 2>&1; echo "pre rc=$?"
tail -n 15 gpurun_out/t_pre.log | cut -c1-300
timeout 300 python tools/k1_bench.py > gpurun_out/k1_bench.log 2>&1; echo "k1 rc=$?"
cat gpurun_out/k1_bench.log | cut -c1-400
if [ "$1" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_u8 -s 3 -c 1 -o gpurun_out/prof_k1 -f python tools/k1_bench.py --sizes 16384 --reps 2 > gpurun_out/ncu_k1.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:preprocess -c 12 --csv --log-file gpurun_out/k1_launches.csv python tools/k1_bench.py --sizes 16384 --reps 2 > /dev/null 2>&1; echo "ncu list rc=$?"
tail -5 gpurun_out/k1_launches.csv | cut -c1-300
fi
