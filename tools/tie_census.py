"""Ghost-pack tie census on the GPU box: free-running ties of the CUDA column
against the oracle with the PRMSSolarGeometry tables from k_soltab (default)
and from the host libm (option soltab_device 0), on DRB, drbx and tile 71 of
the CONUS bench domain (10 years).

    python tools/tie_census.py
"""
import pathlib as pl
import sys

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import numpy as np

import parity
import prms_oracle as po
import scenarios
from pywatershed_b200 import Engine


def census(label, p, f, start, nd):
    ref_st = po.Oracle(p, start=start).soltabs()
    ref, _ = po.Oracle(p, start=start).run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], po.HRU_VARS, ())
    for sdev in (True, False):
        e = Engine(p, start, soltab_device=sdev, channel=False)
        e.select_outputs(e.reg.hru_vars, [])
        got, _ = e.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
        e_soltab = e.soltab()
        e.close()
        names = [v for v in po.HRU_VARS if not v.startswith("channel_")]  # the runs have no channel
        rep = parity.compare_columns(got, ref, names, got["pkwater_equiv"], ref["pkwater_equiv"])
        nbits = 0
        for k in ("pkwater_equiv", "snowmelt", "swrad", "potet", "tcal", "pk_temp"):
            g, r = np.asarray(got[k], dtype=np.float64), ref[k]
            nbits += int(((g.view(np.int64) != r.view(np.int64)) & ~(np.isnan(g) & np.isnan(r))).sum())
        st = e_soltab
        nst = sum(int((np.asarray(st[k]).view(np.int64) != np.asarray(ref_st[k]).view(np.int64)).sum()) for k in st)
        print(f"{label:10s} {nd:5d} days  soltab {'k_soltab (device)' if sdev else 'host libm       '}: "
              f"{len(rep['ties']):4d} ghost-pack ties of {ref['pkwater_equiv'].shape[1]} HRUs, "
              f"{len(rep['mismatches'])} mismatches; values of 6 upper-half variables whose bits differ: {nbits}; "
              f"soltab entries whose bits differ from the oracle's (glibc): {nst} of {sum(np.asarray(v).size for v in st.values())}",
              flush=True)


p, f, start = scenarios.load_drb()
census("drb", p, f, start, 731)
px, fx = scenarios.drbx(p, f)
census("drbx", px, fx, start, 731)
import bench
import torch

spec = bench.workload_spec("conus")
p0, f0, pt, start2 = bench.build_domain(spec)
forc = bench.device_forcings(f0, spec, 0, torch.device("cuda", 0))
t = 71
ft = {k: forc[k][:, t * 765:(t + 1) * 765].cpu().numpy() for k in ("prcp", "tmax", "tmin")}
del forc
census("conus t71", p0, ft, start2, spec["ndays"])
