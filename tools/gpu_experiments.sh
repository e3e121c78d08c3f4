run() { python bench.py --steps 2 --warmup 2 --no-extras --no-cpu "$@" > gpurun_out/bx.json 2>> gpurun_out/bx.err; python -c "
import json,sys;d=json.load(open('gpurun_out/bx.json'));r=d['roofline'];e=d.get('e2e') or {}
print(sys.argv[1:],'value %.4g ms %.2f frac %.3f colshare %.3f chshare %.3f lanes %s e2e %.4g'%(d['value'],d['ms_per_step'],r['frac'],r.get('share_of_step',0),r.get('channel_share_of_step',0),r.get('launches_per_block'),e.get('value',0)))" "$@"; cp gpurun_out/bx.json "gpurun_out/bx_$(echo $@ | tr ' /' '__').json"; }
run --workload ucb
run --workload ucb --overlap 0
run --workload drb
run --workload channel
run --workload channel --block-days 3653
run --workload channel_chained
run --workload channel_chained --block-days 3653
run --outputs nhm_default --no-e2e
run --outputs routing --no-e2e
