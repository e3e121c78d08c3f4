"""ctypes wrapper around the C oracle (oracle/prms_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing in
pywatershed_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import pathlib as pl
import re
import subprocess

import numpy as np

HERE = pl.Path(__file__).resolve().parent
_LIB = {}


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc, no FMA)."""
    so = HERE / "libprms_oracle.so"
    src_m = max((HERE / f).stat().st_mtime for f in ("prms_oracle.c", "prms_oracle.h"))
    if force or not so.exists() or so.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(HERE), "-B"], check=True, capture_output=True)


def _macro_names(macro: str) -> list[str]:
    txt = (HERE / "prms_oracle.h").read_text()
    m = re.search(r"#define " + macro + r"\(X\)((?:.*\\\n)*.*\n)", txt)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    return re.findall(r"X\((\w+)", body)


HRU_F8 = _macro_names("ORC_HRU_F8_PARAMS")
HRU_I4 = _macro_names("ORC_HRU_I4_PARAMS")
MON_F8 = _macro_names("ORC_MON_F8_PARAMS")
MON_I4 = _macro_names("ORC_MON_I4_PARAMS")
SEG_F8 = _macro_names("ORC_SEG_F8_PARAMS")
SEG_I4 = _macro_names("ORC_SEG_I4_PARAMS")
HRU_VARS = _macro_names("ORC_HRU_VARS")
SEG_VARS = _macro_names("ORC_SEG_VARS")
# FlowGraph.get_init_values order (base/flow_graph.py:425-436)
FLOW_GRAPH_VARS = ["outflows", "node_upstream_inflows", "node_outflows", "node_storage_changes",
                   "node_storages", "node_sink_source", "node_negative_sink_source"]
SCALAR_F8 = ["albset_rna", "albset_rnm", "albset_sna", "albset_snm"]
ALL_PARAM_NAMES = (
    HRU_F8 + HRU_I4 + MON_F8 + MON_I4 + SEG_F8 + SEG_I4 + SCALAR_F8
    + ["snarea_curve", "temp_units"]
)

# meta dtypes of the non-float64 public variables (static/metadata/variables.yaml)
INT_VARS = (
    "transp_on", "pptmix", "intcp_form", "intcp_transp_on", "int_alb", "iso",
    "lso", "lst", "mso", "newsnow", "pptmix_nopack",
)
BOOL_VARS = ("iasw",)


def lib(omp: bool = False):
    key = "omp" if omp else "st"
    if key not in _LIB:
        build()
        name = "libprms_oracle_omp.so" if omp else "libprms_oracle.so"
        L = C.CDLL(str(HERE / name))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_f8.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.orc_set_i4.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_init.argtypes = [C.c_void_p]
        L.orc_hru_var_name.restype = C.c_char_p
        L.orc_hru_var_name.argtypes = [C.c_int]
        L.orc_seg_var_name.restype = C.c_char_p
        L.orc_seg_var_name.argtypes = [C.c_int]
        for fn in ("orc_soltab_potsw", "orc_soltab_horad_potsw", "orc_soltab_sunhrs"):
            getattr(L, fn).restype = C.c_void_p
            getattr(L, fn).argtypes = [C.c_void_p]
        L.orc_run.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [
            C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_run_channel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p]
        L.orc_run_flow_graph.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.orc_set_channel_coefficients.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_hru_var.restype = C.c_void_p
        L.orc_hru_var.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_hru_var.restype = C.c_int
        L.orc_set_hru_var.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_seg_var.restype = C.c_void_p
        L.orc_seg_var.argtypes = [C.c_void_p, C.c_int]
        assert [L.orc_hru_var_name(i).decode() for i in range(L.orc_n_hru_vars())] == HRU_VARS
        _LIB[key] = L
    return _LIB[key]


def calendar(start: str | np.datetime64, ndays: int) -> np.ndarray:
    """[ndays,4] int32 (doy, dowy, month, day-of-month), following the
    reference's utils/time_utils.py:25-68."""
    t = np.datetime64(start, "D") + np.arange(ndays)
    year = t.astype("datetime64[Y]")
    month = t.astype("datetime64[M]").astype(int) % 12 + 1
    doy = (t - year.astype("datetime64[D]")).astype(int) + 1
    dom = (t - t.astype("datetime64[M]").astype("datetime64[D]")).astype(int) + 1
    y = year.astype(int) + 1970
    wy_start_year = np.where(month < 10, y - 1, y)
    wy0 = np.array([np.datetime64(f"{yy}-10-01") for yy in wy_start_year])
    dowy = (t - wy0).astype(int) + 1
    return np.ascontiguousarray(np.stack([doy, dowy, month, dom], axis=1), dtype=np.int32)


class Oracle:
    """One domain on the CPU oracle."""

    def __init__(self, params: dict, start: str | np.datetime64 = "1979-01-01",
                 dprst_flag: bool = True, omp: bool = False):
        self.L = lib(omp)
        self.nhru = int(np.asarray(params["hru_area"]).shape[0])
        self.nseg = int(np.asarray(params["tosegment"]).shape[0]) if "tosegment" in params else 0
        ndepl = int(np.asarray(params["snarea_curve"]).size)
        self.h = self.L.orc_create(self.nhru, self.nseg, ndepl)
        self.start = np.datetime64(start, "D")
        for name in HRU_F8 + MON_F8 + SEG_F8 + SCALAR_F8 + ["snarea_curve"]:
            if name in SEG_F8 and self.nseg == 0:
                continue
            self._f8(name, params[name])
        for name in HRU_I4 + MON_I4 + SEG_I4:
            if (name in SEG_I4 or name == "hru_segment") and self.nseg == 0:
                continue
            self._i4(name, params[name])
        cal0 = calendar(self.start, 1)[0]
        self._f8("temp_units", np.array([float(np.asarray(params.get("temp_units", 0)).ravel()[0])]))
        self._f8("start_month", np.array([float(cal0[2])]))
        self._f8("start_doy", np.array([float(cal0[0])]))
        self._f8("dprst_flag", np.array([1.0 if dprst_flag else 0.0]))
        self._check(self.L.orc_init(self.h))
        self.day = 0

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def _f8(self, name, val):
        a = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel())
        self._check(self.L.orc_set_f8(self.h, name.encode(), a.ctypes.data, a.size))

    def _i4(self, name, val):
        a = np.ascontiguousarray(np.asarray(val).ravel().astype(np.int32))
        self._check(self.L.orc_set_i4(self.h, name.encode(), a.ctypes.data, a.size))

    def soltabs(self):
        out = {}
        for nm in ("soltab_potsw", "soltab_horad_potsw", "soltab_sunhrs"):
            p = getattr(self.L, "orc_" + nm)(self.h)
            out[nm] = np.ctypeslib.as_array(
                C.cast(p, C.POINTER(C.c_double)), shape=(366, self.nhru)).copy()
        return out

    def run(self, prcp, tmax, tmin, hru_vars=(), seg_vars=(), budgets=True):
        """Advance len(prcp) days.  Returns (dict var -> [ndays, n] float64,
        budget violation counts [ndays, 6] or None)."""
        prcp = np.ascontiguousarray(prcp, dtype=np.float64)
        tmax = np.ascontiguousarray(tmax, dtype=np.float64)
        tmin = np.ascontiguousarray(tmin, dtype=np.float64)
        nd = prcp.shape[0]
        assert prcp.shape == tmax.shape == tmin.shape == (nd, self.nhru)
        cal = calendar(self.start + self.day, nd)
        hidx = np.array([HRU_VARS.index(v) for v in hru_vars], dtype=np.int32)
        sidx = np.array([SEG_VARS.index(v) for v in seg_vars], dtype=np.int32)
        hout = np.empty((nd, len(hidx), self.nhru))
        sout = np.empty((nd, len(sidx), self.nseg))
        viol = np.zeros((nd, 6), dtype=np.int32) if budgets else None
        self._check(self.L.orc_run(
            self.h, nd, cal.ctypes.data, prcp.ctypes.data, tmax.ctypes.data,
            tmin.ctypes.data, len(hidx), hidx.ctypes.data, hout.ctypes.data,
            len(sidx), sidx.ctypes.data, sout.ctypes.data,
            viol.ctypes.data if budgets else None))
        self.day += nd
        res = {v: hout[:, k, :] for k, v in enumerate(hru_vars)}
        res.update({v: sout[:, k, :] for k, v in enumerate(seg_vars)})
        return res, viol

    def set_hru_vars(self, values: dict):
        """Overwrite the current value of public HRU variables (name -> [nhru])."""
        for k, v in values.items():
            a = np.ascontiguousarray(v, dtype=np.float64)
            assert a.shape == (self.nhru,)
            rc = self.L.orc_set_hru_var(self.h, HRU_VARS.index(k), a.ctypes.data)
            assert rc == 0, k

    def current(self, var):
        """Current value of a public HRU variable ([nhru] copy)."""
        p = self.L.orc_hru_var(self.h, HRU_VARS.index(var))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(self.nhru,)).copy()

    def run_forced(self, prcp, tmax, tmin, series: dict, columns=None):
        """One-step parity: day d is computed from `series`' day d-1 (every
        public HRU variable overwritten before the step) instead of from the
        oracle's own day d-1.  With `columns` (HRU indices) the series are
        [ndays, len(columns)] and only those HRUs are overwritten.  Returns
        var -> [ndays, nhru]."""
        nd = np.asarray(prcp).shape[0]
        out = {k: np.empty((nd, self.nhru)) for k in HRU_VARS}
        for d in range(nd):
            if d > 0:
                if columns is None:
                    vals = {k: np.asarray(series[k][d - 1], dtype=np.float64) for k in HRU_VARS}
                else:
                    vals = {}
                    for k in HRU_VARS:
                        v = self.current(k)
                        v[columns] = np.asarray(series[k][d - 1], dtype=np.float64)
                        vals[k] = v
                self.set_hru_vars(vals)
            res, _ = self.run(prcp[d:d + 1], tmax[d:d + 1], tmin[d:d + 1], HRU_VARS, (), budgets=False)
            for k in HRU_VARS:
                out[k][d] = res[k][0]
        return out

    def run_channel(self, seg_lateral_inflow, seg_vars=SEG_VARS):
        lat = np.ascontiguousarray(seg_lateral_inflow, dtype=np.float64)
        nd = lat.shape[0]
        sidx = np.array([SEG_VARS.index(v) for v in seg_vars], dtype=np.int32)
        sout = np.empty((nd, len(sidx), self.nseg))
        self._check(self.L.orc_run_channel(self.h, nd, lat.ctypes.data, len(sidx),
                                           sidx.ctypes.data, sout.ctypes.data))
        return {v: sout[:, k, :] for k, v in enumerate(seg_vars)}

    def set_channel_coefficients(self, tsi, ts, c0, c1, c2):
        """Override the Muskingum coefficients (a test aid, see prms_oracle.c)."""
        a = [np.ascontiguousarray(tsi, dtype=np.int32)] + [np.ascontiguousarray(v, dtype=np.float64)
                                                           for v in (ts, c0, c1, c2)]
        assert all(v.shape == (self.nseg,) for v in a)
        self._check(self.L.orc_set_channel_coefficients(self.h, *[v.ctypes.data for v in a]))

    def run_flow_graph(self, inflows, kind, obs=None, flow_min=None):
        """FlowGraph over this oracle's segments as nodes (orc_run_flow_graph):
        inflows [ndays, nnodes], kind [nnodes] (0 Muskingum-Mann, 1 pass-through,
        2 observed flow, 3 source / sink), obs [ndays, nnodes]: the ObsIn nodes'
        observations and the SourceSink nodes' requests, flow_min [nnodes].  Returns
        var -> [ndays, nnodes] for FlowGraph's seven variables."""
        lat = np.ascontiguousarray(inflows, dtype=np.float64)
        nd = lat.shape[0]
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        if obs is None:
            if (kind >= 2).any():
                raise ValueError("ObsIn nodes need obs")
            obs = np.zeros((nd, self.nseg))
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        assert lat.shape == obs.shape == (nd, self.nseg) and kind.shape == (self.nseg,)
        out = np.empty((nd, len(FLOW_GRAPH_VARS), self.nseg))
        fm = np.zeros(self.nseg) if flow_min is None else np.ascontiguousarray(flow_min, dtype=np.float64)
        assert fm.shape == (self.nseg,)
        self._check(self.L.orc_run_flow_graph(self.h, nd, lat.ctypes.data, kind.ctypes.data,
                                              obs.ctypes.data, fm.ctypes.data, out.ctypes.data))
        return {v: out[:, k, :] for k, v in enumerate(FLOW_GRAPH_VARS)}

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
