"""Golden vectors for the reference's smallest and longest test domain:
test_data/hru_1 -- ONE HRU draining to ONE segment, 1979-01-01 .. 2019-12-31
(14,975 days, 41 years), full NHM chain.  Pins the oracle over a long run
(drift, ten leap years, forty water-year resets) and the single-unit edge.
From the UNMODIFIED reference (numpy calc_method); build container only.
Writes tests/golden/hru1_inputs.npz and tests/golden/hru1_golden.npz (every
variable on the days listed in "days": the first three and last two years in
full, every 7th day in between).
"""
import pathlib as pl
import sys
import warnings

import numpy as np

HERE = pl.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "tests"))
import gen_golden as gg  # noqa: E402
import prms_oracle as po  # noqa: E402
import ref_harness as rh  # noqa: E402

DOM = pl.Path("/root/reference/test_data/hru_1")


def main():
    pws = rh.import_reference()
    control = pws.Control.load_prms(DOM / "nhm.control", warn_unused_options=False)
    params = pws.parameters.PrmsParameters.load(DOM / "myparam.param")
    from pywatershed.utils.cbh_utils import cbh_files_to_np_dict

    forcings, times = {}, None
    for var in ("prcp", "tmax", "tmin"):
        dd = cbh_files_to_np_dict({var: DOM / f"{var}.cbh"}, params)
        key = [k for k in dd if k not in ("time", "datetime", "nhm_id")]
        assert len(key) == 1, key
        forcings[var] = np.asarray(dd[key[0]], dtype=np.float64)
        times = dd["datetime"] if "datetime" in dd else dd["time"]
    pp = params.parameters
    inputs = {"p_" + n: np.asarray(pp[n]) for n in po.ALL_PARAM_NAMES}
    inputs.update(forcings)
    inputs["start"] = np.array(str(control.start_time.astype("datetime64[D]")))
    np.savez_compressed(gg.GOLDEN / "hru1_inputs.npz", **inputs)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = rh.build_model(pws, control, params, forcings, times, budget_type="error")
        out = rh.run_and_dump(model)
    nd = out["tmaxf"].shape[0]
    days = np.unique(np.concatenate([np.arange(0, 1096), np.arange(1096, nd - 731, 7), np.arange(nd - 731, nd)]))
    z = {"days": days}
    for k, v in out.items():
        z[k if k.startswith("soltab") else "ts_" + k] = v if k.startswith("soltab") else v[days]
    np.savez_compressed(gg.GOLDEN / "hru1_golden.npz", **z)
    print("hru_1:", len(out), "variables,", nd, "days; seg_outflow[-1] =", repr(out["seg_outflow"][-1]),
          "max pkwater_equiv =", repr(out["pkwater_equiv"].max()))


if __name__ == "__main__":
    main()
