# one-GPU tuning lines (run under gpurun): each prints a one-line summary and keeps the JSON
run() { python bench.py --steps 2 --warmup 2 --no-extras --no-cpu "$@" > gpurun_out/bx.json 2>> gpurun_out/bx.err; python -c "
import json,sys;d=json.load(open('gpurun_out/bx.json'));r=d.get('roofline') or {};e=d.get('e2e') or {}
print(sys.argv[1:],'value %.4g ms %.2f frac %.3f colshare %.3f chshare %.3f lanes %s e2e %.4g'%(d['value'],d['ms_per_step'],r.get('frac',0),r.get('share_of_step',0),r.get('channel_share_of_step',0),r.get('launches_per_block'),e.get('value',0)))" "$@"; cp gpurun_out/bx.json "gpurun_out/bx_$(echo $@ | tr ' /' '__').json"; }
for a in "$@"; do run $a; done
