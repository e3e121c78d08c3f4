"""cProfile of pywatershed_b200.Model.run() on the CONUS-scale domain (GPU box)."""
import cProfile
import pstats
import sys
import pathlib as pl
import time

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import torch

import bench
import pywatershed_b200 as pwsb

spec = bench.workload_spec("conus")
ndays = int(sys.argv[1]) if len(sys.argv) > 1 else 731
p0, f0, pt, start = bench.build_domain(spec)
nhru, nseg = pt["hru_area"].shape[0], pt["tosegment"].shape[0]
sub = dict(spec)
sub["ndays"] = ndays
forc = {k: v.float().cpu().numpy() for k, v in bench.device_forcings(f0, sub, 0, torch.device("cuda", 0)).items()}
procs = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff", "PRMSSoilzone",
         "PRMSGroundwater", "PRMSChannel"]


def make():
    t0 = np.datetime64(start, "s")
    control = pwsb.Control(t0, t0 + np.timedelta64(ndays - 1, "D"), np.timedelta64(24, "h"),
                           options={"budget_type": "error", "calc_method": "cuda"})
    coords = {"nhm_id": pt["nhm_id"], "nhm_seg": pt["nhm_seg"]}
    params = pwsb.Parameters(dims={"nhru": nhru, "nsegment": nseg, "nmonth": 12, "ndoy": 366},
                             coords=coords, data_vars={k: v for k, v in pt.items() if k not in coords})
    m = pwsb.Model([getattr(pwsb, n) for n in procs], control=control, parameters=params, find_input_files=False)
    for k in ("prcp", "tmax", "tmin"):
        m.processes["PRMSAtmosphere"].set_input_to_adapter(k, forc[k])
    return m


m = make()
t0 = time.perf_counter(); m._ensure_engine(); print("ensure_engine", time.perf_counter() - t0)
t0 = time.perf_counter(); m.run(finalize=True); print("first run", time.perf_counter() - t0)
m = make()
m._ensure_engine()
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter(); m.run(finalize=True); dt = time.perf_counter() - t0
pr.disable()
print("second run", dt, "s", nhru * ndays / dt, "HRU-days/s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
