# A/B of alternative builds of the library (PWS_B200_LIB) on the default workload
for i in 1 2; do
for lib in default $@; do
  if [ "$lib" = default ]; then unset PWS_B200_LIB; else export PWS_B200_LIB=$lib; fi
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras > /tmp/ab.json 2>/tmp/ab.err
  python - "$lib" <<PY
import json,sys
try:
    d=json.load(open("/tmp/ab.json")); r=d["roofline"]
    print(sys.argv[1] or "default", "value %.4g ms %.2f frac %.3f col_ms %.3f"%(d["value"],d["ms_per_step"],r["frac"],r.get("avg_launch_ms",0)))
except Exception as e:
    print(sys.argv[1], "failed", e, open("/tmp/ab.err").read()[-300:])
PY
done
done
