python -m pytest tests -m gpu -q -s > gpurun_out/t4.log 2>&1; tail -12 gpurun_out/t4.log
run() { python bench.py --steps 2 --warmup 2 --no-extras --no-e2e --no-cpu "$@" > gpurun_out/b4.json 2>> gpurun_out/b4.err; python -c "
import json,sys;d=json.load(open('gpurun_out/b4.json'));r=d['roofline'];print(sys.argv[1:],'value %.4g ms %.2f frac %.3f colshare %.3f chshare %.3f lanes %s'%(d['value'],d['ms_per_step'],r['frac'],r['share_of_step'],r['channel_share_of_step'],r['launches_per_block']))" "$@"; }
run
run --shard-of 8
run --shard-of 8 --overlap 0
run --shard-of 2
run --workload ensemble128
run --workload ensemble128 --lanes 1 --overlap 0
run --workload ensemble
