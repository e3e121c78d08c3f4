"""Which operation of the CUDA column differs from the oracle in the LAST BIT?

One-step (teacher-forced) comparison on the GPU box: the oracle continues from
the GPU's own day d-1 and its day d is compared with the GPU's day d for exact
bit equality, variable by variable in registry (= computation) order.  The first
variables of the chain that show differing HRU-days name the operation.

    python tools/bit_diff.py [drb|drbx] [ndays] [soltab_device 0|1]
"""
import pathlib as pl
import sys

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import numpy as np

import prms_oracle as po
import scenarios
from pywatershed_b200 import Engine

scen = sys.argv[1] if len(sys.argv) > 1 else "drb"
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 731
sdev = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
p, f, start = scenarios.load_drb()
if scen == "drbx":
    p, f = scenarios.drbx(p, f)
e = Engine(p, start, soltab_device=sdev)
e.select_outputs(e.reg.hru_vars, e.reg.seg_vars)
got, _ = e.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
st = e.soltab()
ref_st = po.Oracle(p, start=start).soltabs()
for k in st:
    a, b = np.asarray(st[k]), np.asarray(ref_st[k])
    print(f"soltab {k}: {int((a.view(np.int64) != b.view(np.int64)).sum())} of {a.size} values differ in bits")
ff = {k: v[:nd] for k, v in f.items()}
one = po.Oracle(p, start=start).run_forced(ff["prcp"], ff["tmax"], ff["tmin"], got)
print(f"{scen}, {nd} days, soltab_device={sdev}: HRU-days whose bits differ (one step from the GPU's own day d-1)")
for k in po.HRU_VARS:
    g = np.asarray(got[k], dtype=np.float64)
    r = one[k]
    neq = (g.view(np.int64) != r.view(np.int64)) & ~(np.isnan(g) & np.isnan(r)) & ~((g == 0) & (r == 0))
    n = int(neq.sum())
    if n:
        d, u = np.argwhere(neq)[0]
        rel = np.abs(g - r)[neq] / np.maximum(np.abs(r[neq]), 1e-300)
        print(f"  {k:28s} {n:8d}  first day {d} hru {u}: {g[d, u]!r} vs {r[d, u]!r}  max rel {rel.max():.3g}")
