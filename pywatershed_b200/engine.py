"""Engine: one domain resident on one B200, driven through the C ABI.

This is the only module that talks to libpwsb200; the reference-facing
classes in `processes.py` / `model.py` sit on top of it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def calendar(start, ndays: int) -> np.ndarray:
    """[ndays, 4] int32 (doy, dowy, month, day of month): what the reference's
    Control exposes per step (base/control.py:439-453,
    utils/time_utils.py:25-68)."""
    t = np.datetime64(start, "D") + np.arange(ndays)
    year = t.astype("datetime64[Y]")
    month = t.astype("datetime64[M]").astype(int) % 12 + 1
    doy = (t - year.astype("datetime64[D]")).astype(int) + 1
    dom = (t - t.astype("datetime64[M]").astype("datetime64[D]")).astype(int) + 1
    y = year.astype(int) + 1970
    wy = np.where(month < 10, y - 1, y)
    wy0 = ((wy - 1970) * 12 + 9).astype("datetime64[M]").astype("datetime64[D]")
    dowy = (t - wy0).astype(int) + 1
    return np.ascontiguousarray(np.stack([doy, dowy, month, dom], axis=1), dtype=np.int32)


class Registry:
    """Names / dtypes of the public variables, read from the library."""

    _inst = None

    def __init__(self):
        L = _lib.lib()
        self.hru_vars = [L.pws_hru_var_name(i).decode() for i in range(L.pws_n_hru_vars())]
        self.hru_kind = [L.pws_hru_var_kind(i) for i in range(L.pws_n_hru_vars())]
        self.seg_vars = [L.pws_seg_var_name(i).decode() for i in range(L.pws_n_seg_vars())]
        self.states = [L.pws_hru_state_name(i).decode() for i in range(L.pws_n_hru_state())]
        self.hru_index = {n: i for i, n in enumerate(self.hru_vars)}
        self.seg_index = {n: i for i, n in enumerate(self.seg_vars)}

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def dtype(self, var):
        k = self.hru_kind[self.hru_index[var]]
        return (np.float64, np.int32, np.bool_)[k]


# parameter tables: name -> kind (must match csrc/pws_registry.h; the library
# rejects unknown names, so a drift fails loudly)
SCALAR_OPTIONS = ("albset_rna", "albset_rnm", "albset_sna", "albset_snm")
SEG_PARAMS_F64 = ("mann_n", "seg_depth", "seg_length", "seg_slope", "x_coef", "segment_flow_init")
SEG_PARAMS_I32 = ("segment_type", "tosegment")
HRU_PARAMS_I32 = ("transp_beg", "transp_end", "cov_type", "hru_type", "hru_deplcrv",
                  "melt_force", "melt_look", "soil_type", "hru_segment")
MON_PARAMS_I32 = ("tstorm_mo",)
HRU_PARAMS_F64 = (
    "hru_slope", "hru_aspect", "hru_lat", "hru_area", "hru_in_to_cf", "radj_sppt",
    "radj_wppt", "jh_coef_hru", "transp_tmax", "covden_sum", "covden_win", "srain_intcp",
    "wrain_intcp", "snow_intcp", "potet_sublim", "emis_noppt", "freeh2o_cap", "rad_trncf",
    "snarea_thresh", "den_init", "den_max", "settle_const", "snowpack_init",
    "hru_percent_imperv", "imperv_stor_max", "dprst_frac", "carea_max", "smidx_coef",
    "smidx_exp", "soil_moist_max", "snowinfil_max", "dprst_depth_avg", "dprst_et_coef",
    "dprst_flow_coef", "dprst_frac_init", "dprst_frac_open", "dprst_seep_rate_clos",
    "dprst_seep_rate_open", "sro_to_dprst_imperv", "sro_to_dprst_perv", "va_open_exp",
    "va_clos_exp", "op_flow_thres", "fastcoef_lin", "fastcoef_sq", "pref_flow_den",
    "pref_flow_infil_frac", "sat_threshold", "slowcoef_lin", "slowcoef_sq",
    "soil_moist_init_frac", "soil_rechr_init_frac", "soil_rechr_max_frac", "soil2gw_max",
    "ssr2gw_exp", "ssr2gw_rate", "ssstor_init_frac", "gwflow_coef", "gwsink_coef",
    "gwstor_init")
MON_PARAMS_F64 = (
    "radadj_intcp", "radadj_slope", "tmax_index", "dday_slope", "dday_intcp", "radmax",
    "ppt_rad_adj", "tmax_allsnow", "tmax_allrain_offset", "jh_coef", "tmax_cbh_adj",
    "tmin_cbh_adj", "snow_cbh_adj", "rain_cbh_adj", "adjmix_rain", "cecn_coef")


class Engine:
    """A PRMS/NHM domain on one GPU.

    params: dict name -> ndarray with the reference's parameter names
    (static/metadata/parameters.yaml); `start`: first simulated day.
    """

    def __init__(self, params: dict, start, device: int = 0, dprst_flag: bool = True,
                 adjust_parameters: bool = True, channel: bool = True,
                 channel_chunk: int | None = None):
        self.L = _lib.lib()
        self.reg = Registry.get()
        self.nhru = int(np.asarray(params["hru_area"]).shape[0])
        self.nseg = int(np.asarray(params["tosegment"]).shape[0]) if (
            channel and "tosegment" in params) else 0
        self.start = np.datetime64(start, "D")
        ndepl = int(np.asarray(params["snarea_curve"]).size)
        h = C.c_void_p()
        rc = self.L.pws_create(C.byref(h), device, self.nhru, self.nseg, ndepl)
        if rc != 0:
            raise _lib.PwsLibraryError(self.L.pws_last_error(None).decode())
        self.h = h
        self.device = device
        for name in HRU_PARAMS_F64 + MON_PARAMS_F64 + ("snarea_curve",):
            self._set_f64(name, params[name])
        for name in HRU_PARAMS_I32 + MON_PARAMS_I32:
            if name == "hru_segment" and self.nseg == 0:
                continue
            self._set_i32(name, params[name])
        if self.nseg:
            for name in SEG_PARAMS_F64:
                self._set_f64(name, params[name])
            for name in SEG_PARAMS_I32:
                self._set_i32(name, params[name])
        for name in SCALAR_OPTIONS:
            v = np.ravel(np.asarray(params[name], dtype=np.float64))
            if v.size != 1:
                # the reference compares these un-indexed (prms_snow.py:911-914,
                # 2222): only length-1 arrays work there
                raise ValueError(f"{name} must be scalar-shaped, got {v.size} values")
            self._opt(name, float(v[0]))
        cal0 = calendar(self.start, 1)[0]
        self._opt("temp_units", float(np.ravel(np.asarray(params.get("temp_units", 0)))[0]))
        self._opt("start_month", float(cal0[2]))
        self._opt("start_doy", float(cal0[0]))
        self._opt("dprst_flag", 1.0 if dprst_flag else 0.0)
        self._opt("adjust_parameters", 1.0 if adjust_parameters else 0.0)
        if channel_chunk is not None:
            self._opt("channel_chunk", channel_chunk)
        self._check(self.L.pws_init(self.h))
        self.day = 0
        self.sel_hru: list[str] = []
        self.sel_seg: list[str] = []

    # ---- plumbing
    def _check(self, rc):
        if rc != 0:
            msg = self.L.pws_last_error(self.h).decode()
            if "compute_interflow" in msg:
                raise RuntimeError(msg)  # prms_soilzone.py:1282-1287
            if "CUDA error" in msg:
                raise _lib.PwsLibraryError(msg)
            raise ValueError(msg)

    def _set_f64(self, name, val):
        a = np.ascontiguousarray(np.asarray(val, dtype=np.float64).ravel())
        self._check(self.L.pws_set_param_f64(self.h, name.encode(), a.ctypes.data, a.size))

    def _set_i32(self, name, val):
        a = np.ascontiguousarray(np.asarray(val).ravel().astype(np.int32))
        self._check(self.L.pws_set_param_i32(self.h, name.encode(), a.ctypes.data, a.size))

    def _opt(self, name, value):
        self._check(self.L.pws_set_option(self.h, name.encode(), float(value)))

    def set_block_days(self, n: int):
        self._opt("block_days", n)

    def reset(self):
        """Back to the initial conditions (static tables are kept)."""
        self._check(self.L.pws_reset_state(self.h))
        self.day = 0

    def timer_mark(self, which: int):
        self._check(self.L.pws_timer_mark(self.h, which))

    def timer_elapsed_ms(self) -> float:
        ms = C.c_double()
        self._check(self.L.pws_timer_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def last_launches(self):
        a, b = C.c_int(), C.c_int()
        self.L.pws_last_launches(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def close(self):
        if getattr(self, "h", None):
            self.L.pws_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- outputs
    def select_outputs(self, hru_vars=(), seg_vars=()):
        """Choose which public variables a run materialises (the reference's
        netcdf_output_var_names).  Order is kept."""
        hru_vars, seg_vars = list(hru_vars), list(seg_vars)
        hi = np.array([self.reg.hru_index[v] for v in hru_vars], dtype=np.int32)
        si = np.array([self.reg.seg_index[v] for v in seg_vars], dtype=np.int32)
        self._check(self.L.pws_select_outputs(self.h, len(hi), hi.ctypes.data, len(si), si.ctypes.data))
        self.sel_hru, self.sel_seg = hru_vars, seg_vars
        self.sel_f64 = [v for v in hru_vars if self.reg.dtype(v) is np.float64]
        self.sel_i32 = [v for v in hru_vars if self.reg.dtype(v) is not np.float64]

    def soltab(self):
        out = {}
        for k, name in enumerate(("soltab_potsw", "soltab_horad_potsw", "soltab_sunhrs")):
            a = np.empty((366, self.nhru))
            self._check(self.L.pws_get_soltab(self.h, k, a.ctypes.data))
            out[name] = a
        return out

    # ---- runs
    def run_host(self, prcp, tmax, tmin, budgets: bool = True):
        """Advance len(prcp) days with forcings / outputs in host memory
        (time-blocked, H2D / compute / D2H overlapped).  Returns
        (dict var -> [ndays, n] ndarray with the reference's dtype,
        budget violation counts [ndays, 6] or None)."""
        prcp = np.ascontiguousarray(prcp, dtype=np.float64)
        tmax = np.ascontiguousarray(tmax, dtype=np.float64)
        tmin = np.ascontiguousarray(tmin, dtype=np.float64)
        nd = prcp.shape[0]
        if not (prcp.shape == tmax.shape == tmin.shape == (nd, self.nhru)):
            raise ValueError("forcings must be [ndays, nhru]")
        cal = calendar(self.start + self.day, nd)
        of = np.empty((nd, len(self.sel_f64), self.nhru))
        oi = np.empty((nd, len(self.sel_i32), self.nhru), dtype=np.int32)
        os_ = np.empty((nd, len(self.sel_seg), self.nseg))
        viol = np.zeros((nd, 6), dtype=np.int32) if budgets else None
        self._check(self.L.pws_run_host(
            self.h, nd, cal.ctypes.data, prcp.ctypes.data, tmax.ctypes.data, tmin.ctypes.data,
            of.ctypes.data, oi.ctypes.data, os_.ctypes.data,
            viol.ctypes.data if budgets else None))
        self.day += nd
        res = {v: of[:, k, :] for k, v in enumerate(self.sel_f64)}
        for k, v in enumerate(self.sel_i32):
            a = oi[:, k, :]
            res[v] = a.astype(np.bool_) if self.reg.dtype(v) is np.bool_ else a
        res.update({v: os_[:, k, :] for k, v in enumerate(self.sel_seg)})
        return res, viol

    def run_device(self, nd, d_prcp, d_tmax, d_tmin, d_out_f64=0, d_out_i32=0, d_seg_out=0,
                   viol: np.ndarray | None = None, sync: bool = True):
        """Advance `nd` days with forcings and outputs resident in HBM
        (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        cal = calendar(self.start + self.day, nd)
        self._check(self.L.pws_run_device(
            self.h, nd, cal.ctypes.data, d_prcp, d_tmax, d_tmin, d_out_f64, d_out_i32,
            d_seg_out, viol.ctypes.data if viol is not None else None))
        self.day += nd
        if sync:
            self.sync()

    def run_channel_device(self, nd, d_seg_lat, d_seg_out, sync=True):
        self._check(self.L.pws_run_channel_device(self.h, nd, d_seg_lat, d_seg_out))
        if sync:
            self.sync()

    def sync(self):
        self._check(self.L.pws_sync(self.h))

    # ---- state / instrumentation
    def get_state(self, name):
        a = np.empty(self.nhru)
        self._check(self.L.pws_get_state(self.h, name.encode(), a.ctypes.data))
        return a

    def set_state(self, name, val):
        a = np.ascontiguousarray(val, dtype=np.float64)
        assert a.shape == (self.nhru,)
        self._check(self.L.pws_set_state(self.h, name.encode(), a.ctypes.data))

    def timing(self):
        a, b = C.c_double(), C.c_double()
        self.L.pws_last_timing(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def launch_count(self):
        return int(self.L.pws_launch_count(self.h))

    def kernel_info(self):
        r, l, b = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.pws_column_kernel_info(self.h, C.byref(r), C.byref(l), C.byref(b)))
        return {"regs": r.value, "local_bytes": l.value, "max_blocks_per_sm": b.value}
