"""Deterministic test scenarios derived from the DRB fixture.

Used by oracle/gen_golden.py (which runs the reference on them) and by the
tests (which run the oracle and the CUDA library on them), so both sides see
bit-identical inputs.  Pure index arithmetic, no RNG.
"""
from __future__ import annotations

import pathlib as pl

import numpy as np

GOLDEN = pl.Path(__file__).resolve().parent / "golden"


def load_drb():
    """DRB 2yr parameters + forcings as extracted from the reference's ASCII
    files (pywatershed/data/drb_2yr) by oracle/gen_golden.py."""
    z = np.load(GOLDEN / "drb_inputs.npz")
    params = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    forc = {k: z[k].astype(np.float64) for k in ("prcp", "tmax", "tmin")}
    return params, forc, str(z["start"])


def load_sagehen():
    """sagehen_5yr (128 HRUs, Sierra Nevada, WY1981-85; the reference's
    test_data/sagehen_5yr ASCII files, extracted by oracle/gen_golden_sagehen.py).
    The parameter file has no depression-storage and no channel parameters (the
    reference runs it with the *NoDprst processes and without PRMSChannel); they
    are completed with the stand-ins the product uses for absent processes."""
    from pywatershed_b200.engine import Engine

    z = np.load(GOLDEN / "sagehen_inputs.npz")
    params = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    params = Engine._filled(params)
    params.pop("tosegment", None)
    forc = {k: z[k].astype(np.float64) for k in ("prcp", "tmax", "tmin")}
    return params, forc, str(z["start"])


def load_hru1():
    """hru_1: one HRU, one segment, 1979-2019 (14,975 days); the reference's
    test_data/hru_1 ASCII files, extracted by oracle/gen_golden_hru1.py."""
    z = np.load(GOLDEN / "hru1_inputs.npz")
    params = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    forc = {k: z[k].astype(np.float64) for k in ("prcp", "tmax", "tmin")}
    return params, forc, str(z["start"])


def drbx(params: dict, forc: dict):
    """'drbx': DRB with the branches DRB never takes switched on for a
    pattern of HRUs: SWALE / LAKE / INACTIVE hru_type,
    pref_flow_den > 0, gwsink_coef > 0, clay soil, bare soil and coniferous
    cover, a bounded transpiration season, thunderstorm months, colder and
    wetter forcing on a third of the HRUs (deep packs, melt season), and
    scalar-shaped snow parameters expanded to [nhru]."""
    p = {k: np.array(v, copy=True) for k, v in params.items()}
    f = {k: np.array(v, copy=True) for k, v in forc.items()}
    nhru = p["hru_area"].shape[0]
    i = np.arange(nhru)
    ht = p["hru_type"].copy()
    ht[i % 17 == 3] = 3  # SWALE
    ht[i % 41 == 5] = 2  # LAKE
    ht[i % 53 == 7] = 0  # INACTIVE
    p["hru_type"] = ht
    m = i % 5 == 1
    p["pref_flow_den"] = np.where(m, 0.05 + 0.01 * (i % 7), p["pref_flow_den"])
    p["pref_flow_infil_frac"] = np.where(m, 0.1 + 0.02 * (i % 3), p["pref_flow_infil_frac"])
    p["gwsink_coef"] = np.where(i % 6 == 2, 0.01, p["gwsink_coef"])
    p["soil_type"] = np.where(i % 4 == 0, 3, p["soil_type"])
    p["cov_type"] = np.where(i % 11 == 4, 0, np.where(i % 13 == 6, 4, p["cov_type"]))
    p["transp_beg"] = np.where(i % 2 == 0, 4, p["transp_beg"])
    p["transp_end"] = np.where(i % 2 == 0, 10, p["transp_end"])
    p["transp_tmax"] = np.where(i % 2 == 0, 500.0 + 10.0 * (i % 9), p["transp_tmax"])
    ts = p["tstorm_mo"].copy()
    ts[4:9, :] = (i % 3 == 0).astype(ts.dtype)
    p["tstorm_mo"] = ts
    p["melt_look"] = np.where(i % 3 == 1, 60, p["melt_look"])
    p["melt_force"] = np.where(i % 3 == 1, 120, p["melt_force"])
    p["den_init"] = np.full(nhru, float(np.ravel(p["den_init"])[0])) * (1.0 + 0.1 * (i % 4))
    p["den_max"] = np.full(nhru, float(np.ravel(p["den_max"])[0]))
    p["settle_const"] = np.full(nhru, float(np.ravel(p["settle_const"])[0]))
    p["slowcoef_sq"] = np.where(i % 7 == 0, 0.0, p["slowcoef_sq"])
    p["ssr2gw_exp"] = np.where(i % 8 == 1, 1.5, p["ssr2gw_exp"])
    p["imperv_stor_max"] = np.where(i % 9 == 2, 0.01, p["imperv_stor_max"])
    cold = i % 3 == 2
    f["tmax"] = f["tmax"] - np.where(cold, 14.0, 0.0)[None, :]
    f["tmin"] = f["tmin"] - np.where(cold, 16.0, 0.0)[None, :]
    f["prcp"] = f["prcp"] * np.where(cold, 1.5, 1.0)[None, :]
    return p, f


def tile(params: dict, forc: dict, ntile: int, nhru_keep: int | None = None,
         temp_shift: bool = True):
    """Tile the domain `ntile` times (independent copies of the basin set:
    tosegment / hru_segment offset per tile, SURVEY.md section 10 'tiling
    recipe'), with a per-tile temperature offset and precipitation scale so
    tiles are in different snow regimes."""
    nhru = params["hru_area"].shape[0]
    nseg = params["tosegment"].shape[0]
    p = {}
    for k, v in params.items():
        v = np.asarray(v)
        if k in ("tosegment", "hru_segment", "tosegment_nhm"):
            parts = [np.where(v > 0, v + t * nseg, v) for t in range(ntile)]
            p[k] = np.concatenate(parts)
        elif v.ndim >= 1 and v.shape[-1] in (nhru, nseg) and k not in ("snarea_curve",):
            p[k] = np.concatenate([v] * ntile, axis=-1)
        else:
            p[k] = v.copy()
    f = {}
    for k, v in forc.items():
        parts = []
        for t in range(ntile):
            a = v
            if temp_shift and k in ("tmax", "tmin"):
                a = v + (-15.0 + 25.0 * ((t * 7) % 16) / 15.0)
            elif temp_shift and k == "prcp":
                a = v * (0.6 + 0.1 * ((t * 5) % 9))
            parts.append(a)
        f[k] = np.concatenate(parts, axis=1)
    if nhru_keep is not None:
        for k, v in list(p.items()):
            if v.ndim >= 1 and v.shape[-1] == nhru * ntile and k not in ("snarea_curve",):
                p[k] = np.ascontiguousarray(v[..., :nhru_keep])
        for k in f:
            f[k] = np.ascontiguousarray(f[k][:, :nhru_keep])
    p["nhm_id"] = np.arange(1, p["hru_area"].shape[0] + 1)
    p["nhm_seg"] = np.arange(1, p["tosegment"].shape[0] + 1)
    return p, f


def write_prms_param_file(path, params: dict, nmonth_names=()):
    """PRMS ASCII parameter file (utils/prms5_file_util.py:165-231): a
    dimensions section and one record per parameter; [12, nhru] arrays are
    written HRU-fastest as PRMS does."""
    nhru = int(np.asarray(params["hru_area"]).shape[0])
    nseg = int(np.asarray(params["tosegment"]).shape[0]) if "tosegment" in params else 0
    dims = {"nhru": nhru, "nsegment": nseg, "nmonths": 12, "one": 1,
            "ndeplval": int(np.asarray(params["snarea_curve"]).size)}
    seg_names = ("tosegment", "tosegment_nhm", "seg_length", "seg_slope", "seg_depth", "mann_n", "x_coef",
                 "segment_type", "segment_flow_init", "nhm_seg", "seg_lat", "seg_width", "obsin_segment",
                 "obsout_segment", "seg_cum_area", "seg_elev", "seg_humidity")
    with open(path, "w") as fh:
        fh.write("written by tests/scenarios.py\nVersion: 1.7\n** Dimensions **\n")
        for k, v in dims.items():
            fh.write(f"####\n{k}\n{v}\n")
        fh.write("** Parameters **\n")
        for k, v in params.items():
            v = np.asarray(v)
            if v.dtype.kind not in "fiu":
                continue
            if v.ndim == 2:
                dn, flat = ["nhru", "nmonths"], v.ravel()  # [12, nhru] C order == (nhru, nmonths) F order
            elif k == "snarea_curve":
                dn, flat = ["ndeplval"], v.ravel()
            elif v.size == 1:
                dn, flat = ["one"], v.ravel()
            elif k in seg_names or (v.shape[0] == nseg and nseg != nhru):
                dn, flat = ["nsegment"], v
            else:
                dn, flat = ["nhru"], v
            tcode = 1 if v.dtype.kind in "iu" else 3
            fh.write(f"####\n{k}\n{len(dn)}\n" + "\n".join(dn) + f"\n{flat.size}\n{tcode}\n")
            fh.write("\n".join(repr(int(x)) if tcode == 1 else repr(float(x)) for x in flat) + "\n")


def flow_graph_params(params: dict, to_graph_index) -> dict:
    """The parameter set of a FlowGraph whose first nsegment nodes are the
    domain's PRMSChannel segments (base/flow_graph.py; the layout of
    prms_channel_flow_graph.py:_build_flow_graph_inputs): nodes are segments,
    tosegment = to_graph_index + 1, and the nodes behind the segments get
    stand-in channel parameters (their kind says they are not Muskingum nodes)."""
    tgi = np.asarray(to_graph_index, dtype=np.int64)
    nn = tgi.shape[0]
    nseg = np.asarray(params["tosegment"]).shape[0]
    p = {}
    for k, v in params.items():
        v = np.asarray(v)
        if k in ("tosegment", "tosegment_nhm"):
            p[k] = (tgi + 1).astype(v.dtype)
        elif v.ndim == 1 and v.shape[0] == nseg and k not in ("hru_segment",) and nseg != np.asarray(
                params["hru_area"]).shape[0]:
            fill = {"mann_n": 0.04, "seg_depth": 1.0, "seg_length": 1000.0, "seg_slope": 0.001,
                    "x_coef": 0.2}.get(k, 0)
            p[k] = np.concatenate([v, np.full(nn - nseg, fill, dtype=v.dtype)])
        else:
            p[k] = v.copy()
    p["nhm_seg"] = np.arange(1, nn + 1)
    return p
