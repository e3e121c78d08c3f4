"""Generate tests/golden/ from the UNMODIFIED reference (numpy calc_method).

Runs only in the build container (needs /root/reference).  Usage:
    python oracle/gen_golden.py
Writes
  tests/golden/drb_inputs.npz   parameters ("p_<name>") + CBH forcings of
                                pywatershed/data/drb_2yr (ASCII inputs)
  tests/golden/drb_golden.npz   reference outputs, full NHM chain, 731 days
  tests/golden/drbx_golden.npz  same for the branch-coverage variant
                                tests/scenarios.py:drbx
Outputs are sampled to keep the fixtures small: all days for a stratified
sample of HRUs / segments, plus every unit on a few snapshot days, plus
per-day domain sums of every variable.
"""
from __future__ import annotations

import pathlib as pl
import sys
import warnings

import numpy as np

HERE = pl.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "tests"))
import prms_oracle as po  # noqa: E402
import ref_harness as rh  # noqa: E402
import scenarios  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
SNAP_DAYS = (0, 120, 420, 730)


def hru_sample(nhru):
    # every 24th HRU plus the first of each drbx pattern class
    s = set(range(0, nhru, 255))
    s.update([2, 3, 4, 5, 6, 7, 245, 584])
    return np.array(sorted(x for x in s if x < nhru))


def seg_sample(tosegment):
    nseg = len(tosegment)
    s = set(range(0, nseg, 24))
    s.update(np.where(tosegment == 0)[0].tolist())  # outlets
    return np.array(sorted(s))


def pack(out, tosegment):
    nhru = out["tmaxf"].shape[1]
    hs, ss = hru_sample(nhru), seg_sample(tosegment)
    z = {"hru_sample": hs, "seg_sample": ss, "snap_days": np.array(SNAP_DAYS)}
    for k, v in out.items():
        if k.startswith("soltab"):
            z[k] = v[:, hs]
            continue
        samp = ss if k in po.SEG_VARS else hs
        z["ts_" + k] = v[:, samp]
        z["snap_" + k] = v[list(SNAP_DAYS), :]
        z["sum_" + k] = v.astype(np.float64).sum(axis=1)
    return z


def run_reference(pws, control, params, forcings, times, budget_type):
    model = rh.build_model(pws, control, params, forcings, times, budget_type=budget_type)
    return rh.run_and_dump(model)


def main():
    pws = rh.import_reference()
    control, params, forcings, times = rh.load_drb(pws)
    pp = params.parameters
    inputs = {"p_" + n: np.asarray(pp[n]) for n in po.ALL_PARAM_NAMES}
    inputs["p_nhm_id"] = np.asarray(params.coords["nhm_id"])
    inputs["p_nhm_seg"] = np.asarray(params.coords["nhm_seg"])
    inputs.update(forcings)
    inputs["start"] = np.array(str(control.start_time.astype("datetime64[D]")))
    np.savez_compressed(GOLDEN / "drb_inputs.npz", **inputs)

    out = run_reference(pws, control, params, forcings, times, "error")
    np.savez_compressed(GOLDEN / "drb_golden.npz", **pack(out, np.asarray(pp["tosegment"])))
    print("drb: sum(seg_outflow[-1]) =", repr(out["seg_outflow"][-1].sum()),
          "sum(pkwater_equiv[-1]) =", repr(out["pkwater_equiv"][-1].sum()))

    # ---- drbx: same domain, branch-coverage parameter/forcing variant ----
    p0, f0, _ = scenarios.load_drb()
    px, fx = scenarios.drbx(p0, f0)
    control2, params2, _, _ = rh.load_drb(pws)
    dd = params2.to_dd()
    for k, v in px.items():
        if k in dd.data_vars:
            old = dd.data_vars[k]
            dd.data_vars[k] = np.asarray(v).astype(old.dtype).reshape(
                old.shape if np.asarray(v).size == old.size else np.asarray(v).shape)
            if dd.data_vars[k].shape != old.shape:
                dd.metadata[k]["dims"] = ("nhru",)
    paramsx = pws.parameters.PrmsParameters(
        dims=dd.dims, coords=dd.coords, data_vars=dd.data_vars,
        metadata=dd.metadata, validate=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        outx = run_reference(pws, control2, paramsx, fx, times, None)
    np.savez_compressed(GOLDEN / "drbx_golden.npz", **pack(outx, np.asarray(pp["tosegment"])))
    print("drbx: sum(seg_outflow[-1]) =", repr(outx["seg_outflow"][-1].sum()),
          "sum(pkwater_equiv[-1]) =", repr(outx["pkwater_equiv"][-1].sum()))
    np.savez("/tmp/refdump/drbx_full.npz", **outx)


if __name__ == "__main__":
    main()
