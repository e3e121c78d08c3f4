# N-GPU measurements: the multi-GPU tests, then the default bench line under torchrun
N=$1
python -m pytest tests -m gpu -q -k "shard or two_gpu or nccl or multi" > gpurun_out/r02_gpu_tests_${N}gpu.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r02_bench_${N}gpu_final.json 2> gpurun_out/r02_bench_${N}gpu_final.err; tail -c 300 gpurun_out/r02_bench_${N}gpu_final.json
