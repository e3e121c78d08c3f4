"""Generate tests/golden/drbx_ties.npz from the UNMODIFIED reference.

drbx is the fixture with the SWALE / LAKE / preferential-flow / clay branches,
and the one where the oracle and the reference part ways on a few HRUs after a
ghost-pack tie (tests/parity.py).  The main drbx fixture only holds domain sums
and a sample of HRUs, so those HRUs used to go unchecked.  This fixture adds
  tied_hrus          the HRUs whose free-running oracle leaves rtol 1e-9
  tied_ts_<var>      the reference's full series of every public variable on
                     exactly those HRUs  -> one-step parity of the oracle
                     against the reference where the free-running check is blind
  sumc_<var>         per-day sums of every variable over all OTHER HRUs
                     -> the checksum of checksums on the tie-free part
Runs only in the build container (needs /root/reference)."""
from __future__ import annotations

import pathlib as pl
import sys
import warnings

import numpy as np

HERE = pl.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "tests"))
import parity  # noqa: E402
import prms_oracle as po  # noqa: E402
import ref_harness as rh  # noqa: E402
import scenarios  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def main():
    pws = rh.import_reference()
    control, params, forcings, times = rh.load_drb(pws)
    p0, f0, start = scenarios.load_drb()
    px, fx = scenarios.drbx(p0, f0)
    dd = params.to_dd()
    for k, v in px.items():
        if k in dd.data_vars:
            old = dd.data_vars[k]
            dd.data_vars[k] = np.asarray(v).astype(old.dtype).reshape(
                old.shape if np.asarray(v).size == old.size else np.asarray(v).shape)
            if dd.data_vars[k].shape != old.shape:
                dd.metadata[k]["dims"] = ("nhru",)
    paramsx = pws.parameters.PrmsParameters(dims=dd.dims, coords=dd.coords, data_vars=dd.data_vars,
                                            metadata=dd.metadata, validate=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = rh.build_model(pws, control, paramsx, fx, times, budget_type=None)
        ref = rh.run_and_dump(model)
    nd = ref["pkwater_equiv"].shape[0]
    o = po.Oracle(px, start=start)
    got, _ = o.run(fx["prcp"][:nd], fx["tmax"][:nd], fx["tmin"][:nd], po.HRU_VARS, ())
    refd = {k: np.asarray(ref[k], dtype=np.float64) for k in po.HRU_VARS}
    rep = parity.compare_columns(got, refd, po.HRU_VARS, got["pkwater_equiv"], refd["pkwater_equiv"])
    assert rep["mismatches"] == [], rep["mismatches"][:5]
    tied = np.array(sorted(t[0] for t in rep["ties"]), dtype=np.int64)
    clean = np.ones(refd["pkwater_equiv"].shape[1], dtype=bool)
    clean[tied] = False
    z = {"tied_hrus": tied, "tie_days": np.array([t[1] for t in sorted(rep["ties"])], dtype=np.int64)}
    for k in po.HRU_VARS:
        z["tied_ts_" + k] = refd[k][:, tied]
        z["sumc_" + k] = refd[k][:, clean].sum(axis=1)
    np.savez_compressed(GOLDEN / "drbx_ties.npz", **z)
    print("drbx: tied HRUs", tied.tolist(), "on days", z["tie_days"].tolist())


if __name__ == "__main__":
    main()
