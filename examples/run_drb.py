"""Run the Delaware River Basin 2-year NHM configuration (765 HRUs, 456
segments, 731 days) through the pywatershed-style API on one B200 and write
the reference's netCDF output files.

    python examples/run_drb.py [output_dir]

Inputs come from tests/golden/drb_inputs.npz (the reference's drb_2yr ASCII
parameter and CBH files, extracted once by oracle/gen_golden.py).  With the
reference's own files the same model is

    control = pws.Control.load_prms("nhm.control")
    params = pws.PrmsParameters.load("myparam.param")
    pws.Model([...], control=control, parameters=params).run(netcdf_dir="out")
"""
import pathlib as pl
import sys
import time

import numpy as np

ROOT = pl.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import pywatershed_b200 as pws  # noqa: E402


def main():
    out_dir = pl.Path(sys.argv[1]) if len(sys.argv) > 1 else None
    z = np.load(ROOT / "tests" / "golden" / "drb_inputs.npz")
    p = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    nhru, nseg = p["hru_area"].shape[0], p["tosegment"].shape[0]
    params = pws.Parameters(
        dims={"nhru": nhru, "nsegment": nseg, "nmonth": 12, "ndoy": 366},
        coords={"nhm_id": p.pop("nhm_id"), "nhm_seg": p.pop("nhm_seg")}, data_vars=p)
    start = np.datetime64(str(z["start"]), "s")
    nd = z["prcp"].shape[0]
    control = pws.Control(start, start + np.timedelta64(nd - 1, "D"), np.timedelta64(24, "h"),
                          options={"budget_type": "error", "calc_method": "cuda"})
    if out_dir is not None:
        control.options["netcdf_output_dir"] = out_dir
        control.options["netcdf_output_var_names"] = ["pkwater_equiv", "soil_moist", "seg_outflow"]
    model = pws.Model(
        [pws.PRMSSolarGeometry, pws.PRMSAtmosphere, pws.PRMSCanopy, pws.PRMSSnow, pws.PRMSRunoff,
         pws.PRMSSoilzone, pws.PRMSGroundwater, pws.PRMSChannel],
        control=control, parameters=params, find_input_files=False)
    for k in ("prcp", "tmax", "tmin"):
        model.processes["PRMSAtmosphere"].set_input_to_adapter(k, z[k])
    t0 = time.perf_counter()
    model.run(netcdf_dir=out_dir, finalize=True)
    dt = time.perf_counter() - t0
    q = model.processes["PRMSChannel"]["seg_outflow"]
    swe = model.processes["PRMSSnow"]["pkwater_equiv"]
    print(f"{nd} days x {nhru} HRUs in {dt:.2f} s ({nd * nhru / dt:.3g} HRU-days/s incl. Python and output)")
    print(f"last day: sum(seg_outflow) = {q.sum()!r} cfs (reference: 214752.70363439058), "
          f"sum(pkwater_equiv) = {swe.sum()!r} in (reference: 186.00889550764475)")
    if out_dir is not None:
        print("wrote", sorted(f.name for f in pl.Path(out_dir).glob("*.nc")))


if __name__ == "__main__":
    main()
