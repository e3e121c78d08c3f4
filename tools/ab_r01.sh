for i in 1 2 3; do
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras > gpurun_out/ab_r02_$i.json 2>/dev/null
(cd _r01 && python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > ../gpurun_out/ab_r01_$i.json 2>/dev/null)
done
python - <<PY
import json
for i in (1,2,3):
    for v in ("r02","r01"):
        d=json.load(open(f"gpurun_out/ab_{v}_{i}.json")); r=d["roofline"]
        print(v, i, "value %.4g ms %.2f frac %.3f col_ms %.3f"%(d["value"],d["ms_per_step"],r["frac"],r.get("avg_launch_ms",0)))
PY
