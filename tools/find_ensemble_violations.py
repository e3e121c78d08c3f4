"""Diagnostic: which members of the bench ensemble have budget verdicts on the
GPU, and does the oracle agree for those members?  (GPU box.)"""
import sys
import pathlib as pl

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import numpy as np

import bench
import prms_oracle as po
import scenarios
from pywatershed_b200 import Engine

p, f, start = scenarios.load_drb()
nd = 731
fb = {k: np.ascontiguousarray(v[:nd]) for k, v in f.items()}
bad = []
CH = 64
for m0 in range(0, 1024, CH):
    mem = np.arange(m0, m0 + CH)
    e = Engine(bench.ensemble_parameters(p, mem), start, forcing_period=765)
    e.select_outputs([], ["seg_outflow"])
    _, viol = e.run_host(fb["prcp"], fb["tmax"], fb["tmin"])
    e.close()
    if viol.sum():
        print("chunk", m0, "violations", viol.sum(axis=0), "days", np.argwhere(viol.sum(axis=1))[:, 0][:10], flush=True)
        for m in mem:
            e = Engine(bench.ensemble_parameters(p, [m]), start, forcing_period=765)
            e.select_outputs([], ["seg_outflow"])
            _, v1 = e.run_host(fb["prcp"], fb["tmax"], fb["tmin"])
            e.close()
            if v1.sum():
                bad.append(int(m))
                o = po.Oracle(bench.ensemble_parameters(p, [m]), start=start)
                _, ov = o.run(fb["prcp"], fb["tmax"], fb["tmin"], (), ())
                o.close()
                print("  member", m, "gpu", np.argwhere(v1).tolist()[:6], v1[v1 > 0][:6], "oracle",
                      np.argwhere(ov).tolist()[:6], ov[ov > 0][:6], flush=True)
print("members with violations:", bad)
