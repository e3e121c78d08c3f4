# final one-GPU measurements of the round: tests, the default bench line, the reference arm,
# other workloads, the launch list and ncu captures of the two kernels
python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 400 gpurun_out/r02_bench_final.json
python bench.py --impl reference > gpurun_out/r02_bench_reference_final.json 2>> gpurun_out/r02_bench_final.err
bash tools/gpu_experiments.sh "--workload channel" "--workload channel --block-days 3653" "--workload channel_chained" "--workload channel_chained --block-days 3653" "--workload channel_chained --block-days 1024" "--workload ucb" "--workload drb" "--outputs nhm_default --no-e2e" "--outputs nhm_default --no-e2e --lib-opt nhm_sink=0" "--outputs routing --no-e2e" "--no-e2e --layout rows" 2>&1 | tee gpurun_out/r02_final_experiments.log
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_route_chunks -s 3 -c 1 -o gpurun_out/prof_route -f $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hru_column -s 3 -c 1 -o gpurun_out/prof_col -f $CMD > gpurun_out/ncu3.log 2>&1
