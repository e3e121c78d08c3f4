"""Golden vectors on a second, snow-dominated domain: the reference's
test_data/sagehen_5yr (128 HRUs in the Sierra Nevada, 1980-10-01 .. 1985-09-30,
two snow depletion curves, no depression storage, no channel parameters), run
as its own sagehen_no_cascades_model.yaml prescribes: SolarGeometry, Atmosphere,
Canopy, Snow, RunoffNoDprst, SoilzoneNoDprst, GroundwaterNoDprst.  From the
UNMODIFIED reference (numpy calc_method); build container only.  Writes
  tests/golden/sagehen_inputs.npz  parameters ("p_<name>") + CBH forcings
  tests/golden/sagehen_golden.npz  every public variable: all days for a
                                   sample of HRUs, all HRUs on snapshot days,
                                   per-day domain sums
"""
import pathlib as pl
import sys
import warnings

import numpy as np

HERE = pl.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "tests"))
import gen_golden as gg  # noqa: E402
import ref_harness as rh  # noqa: E402

DOM = pl.Path("/root/reference/test_data/sagehen_5yr")
PROCS = ("PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoffNoDprst",
         "PRMSSoilzoneNoDprst", "PRMSGroundwaterNoDprst")
SNAP_DAYS = (0, 150, 240, 900, 1825)


def main():
    pws = rh.import_reference()
    control = pws.Control.load_prms(DOM / "sagehen_no_cascades.control", warn_unused_options=False)
    control.options["dprst_flag"] = False
    params = pws.parameters.PrmsParameters.load(DOM / "myparam.param")
    from pywatershed.utils.cbh_utils import cbh_files_to_np_dict

    forcings, times = {}, None
    for var in ("prcp", "tmax", "tmin"):  # the files name them hru_ppt / tmaxf / tminf
        dd = cbh_files_to_np_dict({var: DOM / f"{var}.cbh"}, params)
        key = [k for k in dd if k not in ("time", "datetime", "nhm_id")]
        assert len(key) == 1, key
        forcings[var] = np.asarray(dd[key[0]], dtype=np.float64)
        times = dd["datetime"] if "datetime" in dd else dd["time"]
    pp = params.parameters
    inputs = {"p_" + n: np.asarray(v) for n, v in pp.items()
              if np.asarray(v).dtype.kind in "fiu" and np.asarray(v).size <= 12 * 128}
    inputs.update(forcings)
    inputs["start"] = np.array(str(control.start_time.astype("datetime64[D]")))
    np.savez_compressed(gg.GOLDEN / "sagehen_inputs.npz", **inputs)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = rh.build_model(pws, control, params, forcings, times, process_names=PROCS,
                               budget_type="error")
        out = rh.run_and_dump(model)
    nhru = out["tmaxf"].shape[1]
    elev = np.asarray(pp["hru_elev"]) if "hru_elev" in pp else np.arange(nhru)
    hs = np.sort(np.argsort(elev)[np.linspace(0, nhru - 1, 6).astype(int)])  # low .. high HRUs
    z = {"hru_sample": hs, "snap_days": np.array(SNAP_DAYS)}
    for k, v in out.items():
        if k.startswith("soltab"):
            z[k] = v[:, hs]
            continue
        z["ts_" + k] = v[:, hs]
        z["snap_" + k] = v[list(SNAP_DAYS), :]
        z["sum_" + k] = v.astype(np.float64).sum(axis=1)
    np.savez_compressed(gg.GOLDEN / "sagehen_golden.npz", **z)
    print("sagehen:", len(out), "variables,", out["tmaxf"].shape, "max pkwater_equiv =",
          repr(out["pkwater_equiv"].max()), "sum(gwres_flow_vol[-1]) =", repr(out["gwres_flow_vol"][-1].sum()))


if __name__ == "__main__":
    main()
