python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
bash tools/gpu_experiments.sh "--workload channel" "--no-e2e" "--no-e2e --shard-of 8" "--outputs routing" 2>&1 | tee gpurun_out/exp3.log
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_route_chunks -s 3 -c 1 -o gpurun_out/prof_route -f $CMD > gpurun_out/ncu2.log 2>&1
