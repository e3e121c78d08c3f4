"""Experiment: how much does grouping HRUs by a static regime key (less warp
divergence) buy the column kernel?  The domain's HRUs are permuted on the host
before the engine sees them (the library treats HRU order as given)."""
import sys
import pathlib as pl

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import torch

import bench
from pywatershed_b200 import Engine

spec = bench.workload_spec("conus")
p0, f0, pt, start = bench.build_domain(spec)
dev = torch.device("cuda", 0)
nd, T = 1024, 256
sub = dict(spec)
sub["ndays"] = nd
forc0 = bench.device_forcings(f0, sub, 0, dev)
nhru = pt["hru_area"].shape[0]


def key_static(p):
    k = np.zeros(nhru, dtype=np.int64)
    k = k * 4 + np.asarray(p["hru_type"])
    k = k * 2 + (np.asarray(p["dprst_frac"]) > 0)
    k = k * 2 + (np.asarray(p["hru_percent_imperv"]) > 0)
    k = k * 8 + np.asarray(p["cov_type"])
    k = k * 4 + np.asarray(p["soil_type"])
    k = k * 2 + (np.asarray(p["pref_flow_den"]) > 0)
    return k


def key_cold(p):
    # colder HRUs keep a pack longer: the monthly temperature adjustment is the per-HRU part of the climate
    return np.asarray(p["tmax_cbh_adj"]).mean(axis=0)


def permutation(kind, window):
    idx = np.arange(nhru)
    if kind == "none":
        return idx
    ks = key_static(pt)
    kc = key_cold(pt)
    out = []
    for w0 in range(0, nhru, window):
        sl = idx[w0:w0 + window]
        if kind == "static":
            o = np.lexsort((sl, ks[sl]))
        elif kind == "static+cold":
            o = np.lexsort((sl, kc[sl], ks[sl]))
        elif kind == "cold":
            o = np.lexsort((sl, kc[sl]))
        elif kind == "dprst":
            o = np.lexsort((sl, (np.asarray(pt["dprst_frac"]) > 0)[sl]))
        out.append(sl[o])
    return np.concatenate(out)


def run(perm):
    q = {}
    for k, v in pt.items():
        v = np.asarray(v)
        q[k] = np.ascontiguousarray(v[..., perm]) if (v.ndim >= 1 and v.shape[-1] == nhru and k not in (
            "snarea_curve",) and k not in ("tosegment", "mann_n", "seg_depth", "seg_length", "seg_slope", "x_coef",
                                           "segment_flow_init", "segment_type", "nhm_seg", "tosegment_nhm",
                                           "seg_lat", "seg_width", "obsin_segment", "obsout_segment", "seg_cum_area",
                                           "seg_elev", "seg_humidity")) else v
    pdev = torch.from_numpy(perm).to(dev)
    forc = {k: v[:, pdev].contiguous() for k, v in forc0.items()}
    eng = Engine(q, start, device=0)
    eng.select_outputs(eng.reg.hru_vars, eng.reg.seg_vars)
    eng.set_tiled_outputs(True)
    onhru = eng.output_tile()[1]
    of = torch.empty((T, len(eng.sel_f64), onhru), dtype=torch.float64, device=dev)
    oi = torch.empty((T, len(eng.sel_i32), onhru), dtype=torch.int32, device=dev)
    osg = torch.empty((T, 5, eng.nseg), dtype=torch.float64, device=dev)
    res = []
    for it in range(3):
        eng.reset()
        torch.cuda.synchronize()
        d = 0
        while d < nd:
            n = min(T, nd - d)
            o = d * nhru * 8
            eng.run_device(n, forc["prcp"].data_ptr() + o, forc["tmax"].data_ptr() + o, forc["tmin"].data_ptr() + o,
                           of.data_ptr(), oi.data_ptr(), osg.data_ptr(), sync=False)
            d += n
        eng.sync()
        res.append(eng.timing()[0])
    chk = float(osg[-1, 3].sum())
    eng.close()
    del forc, of, oi, osg
    torch.cuda.empty_cache()
    return min(res[1:]), chk


nseg_names = None
base = None
for kind, window in (("none", nhru), ("static", 256), ("static", 768), ("static", 4096), ("static+cold", 768),
                     ("cold", 768), ("dprst", 256), ("static", nhru)):
    ms, chk = run(permutation(kind, window))
    base = base or ms
    print(f"{kind:12s} window {window:7d}: column {ms:8.2f} ms per {nd} days  ({base / ms:5.3f}x)  checksum {chk:.6f}",
          flush=True)
