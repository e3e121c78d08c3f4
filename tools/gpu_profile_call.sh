# one gpurun call: GPU tests, small-grid experiments, then the launch list and ncu captures
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
bash tools/gpu_experiments.sh "--no-e2e --shard-of 8" "--workload ucb --no-e2e" "--workload drb --no-e2e" 2>&1 | tee gpurun_out/exp2.log
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_route_chunks -s 3 -c 1 -o gpurun_out/prof_route -f $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hru_column -s 3 -c 1 -o gpurun_out/prof_col -f $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-300; ls -la gpurun_out/
