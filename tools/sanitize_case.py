"""A small case for compute-sanitizer: every kernel of the library on a domain
with several column tiles, two lanes, routing beside the next block, outlet
trees cut into chunks (cross-chunk edges through HBM), a shared-forcing
ensemble, accumulations and the PRMSEt kernel; checked against the oracle."""
import sys
import pathlib as pl

ROOT = pl.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import numpy as np

import prms_oracle as po
import scenarios
from pywatershed_b200 import Engine
from pywatershed_b200.engine import EtEngine

p, f, start = scenarios.load_drb()
nd = 9
pt, ft = scenarios.tile(p, {k: v[:nd] for k, v in f.items()}, 2)
e = Engine(pt, start, channel_chunk=100, lanes=2)
e._opt("overlap_routing", 1)
e.select_outputs(e.reg.hru_vars, e.reg.seg_vars)
e.select_accumulations(["net_rain", "hru_actet"], ["seg_stor_change"])
e.set_block_days(4)
got, viol = e.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
acc = e.accumulations()
e.close()
ref, oviol = po.Oracle(pt, start=start).run(ft["prcp"], ft["tmax"], ft["tmin"], po.HRU_VARS, po.SEG_VARS)
for k in ("pkwater_equiv", "soil_moist", "sroff_vol", "seg_outflow"):
    np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), ref[k], rtol=1e-9, atol=1e-10, err_msg=k)
np.testing.assert_array_equal(viol, oviol)
np.testing.assert_allclose(acc["hru_actet"], ref["hru_actet"].sum(axis=0), rtol=1e-12)
# tiled layout + shared-forcing ensemble (the kPeriod kernel variant, wrap-around TMA rows)
pe, _ = scenarios.tile({k: (np.asarray(v)[..., :764] if np.asarray(v).ndim and np.asarray(v).shape[-1] == 765 else v)
                        for k, v in p.items()}, {k: np.zeros((2, 764)) for k in ("prcp", "tmax", "tmin")}, 3,
                       temp_shift=False)
e = Engine(pe, start, forcing_period=764)
e.select_outputs(e.reg.hru_vars, e.reg.seg_vars)
e.set_tiled_outputs(True)
fb = {k: np.ascontiguousarray(v[:nd, :764]) for k, v in f.items()}
til, _ = e.run_host(fb["prcp"], fb["tmax"], fb["tmin"], pinned_outputs=True)
assert np.isfinite(np.asarray(til["soil_moist"])).all()
e.close()
et = EtEngine(p, start)
out = et.run({k: ref[k][:, :765] for k in EtEngine.INPUTS})
assert np.isfinite(out["hru_actet"]).all()
print("sanitize case ok")
