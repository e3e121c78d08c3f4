"""Golden vectors for the reference's *NoDprst process variants
(PRMSRunoffNoDprst, PRMSSoilzoneNoDprst, PRMSGroundwaterNoDprst: the same
physics with dprst_flag=False, pywatershed/hydrology/prms_*_no_dprst.py),
from the UNMODIFIED reference (numpy calc_method) on drb_2yr.  Build container
only.  Writes tests/golden/drb_nodprst_golden.npz (packed like drb_golden.npz).
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

NODPRST = ("PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoffNoDprst",
           "PRMSSoilzoneNoDprst", "PRMSGroundwaterNoDprst", "PRMSChannel")


def main():
    pws = rh.import_reference()
    control, params, forcings, times = rh.load_drb(pws)
    control.options["dprst_flag"] = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = rh.build_model(pws, control, params, forcings, times, process_names=NODPRST,
                               budget_type="error")
        out = rh.run_and_dump(model, n_days=400)
    tos = np.asarray(params.parameters["tosegment"])
    z = {}
    nhru = out["tmaxf"].shape[1]
    hs, ss = gg.hru_sample(nhru), gg.seg_sample(tos)
    z["hru_sample"], z["seg_sample"] = hs, ss
    seg = ("channel_outflow_vol", "seg_lateral_inflow", "seg_upstream_inflow", "seg_outflow",
           "seg_stor_change")
    for k, v in out.items():
        if k.startswith("soltab"):
            continue
        z["ts_" + k] = v[:, ss if k in seg else hs]
        z["sum_" + k] = v.astype(np.float64).sum(axis=1)
    np.savez_compressed(gg.GOLDEN / "drb_nodprst_golden.npz", **z)
    print("nodprst:", len(out), "variables; sum(seg_outflow[-1]) =", repr(out["seg_outflow"][-1].sum()))
    print(sorted(k for k in out if "dprst" in k))


if __name__ == "__main__":
    main()
