for d in _r01 . _r01 .; do
  (cd $d && python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $( [ "$d" != "_r01" ] && echo --no-extras ) > /tmp/ab.json 2>/tmp/ab.err; python - "$d" <<PY
import json,sys
try:
    d=json.load(open("/tmp/ab.json")); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms %.2f frac %.3f col_ms %.3f"%(d["value"],d["ms_per_step"],r["frac"],r.get("avg_launch_ms",0)))
except Exception as e:
    print(sys.argv[1], "failed", e, open("/tmp/ab.err").read()[-300:])
PY
)
done
