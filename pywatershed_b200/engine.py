"""Engine: one domain resident on one B200, driven through the C ABI.

This is the only module that talks to libpwsb200; the reference-facing
classes in `processes.py` / `model.py` sit on top of it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def untile(out, slot: int, nhru: int):
    """Variable number `slot` of a tiled result buffer [nd, tiles, n_vars,
    tile_units] (include/pws_b200.h: pws_output_tile) as [nd, nhru]: the tiles'
    rows of that variable side by side, the padding of the last tile cut off."""
    v = out[:, :, slot, :]
    return v.reshape(v.shape[0], -1)[:, :nhru]


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
        self.inputs = [L.pws_input_name(i).decode() for i in range(L.pws_n_inputs())]
        self.input_index = {n: i for i, n in enumerate(self.inputs)}
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


# Stand-in values for parameters of processes that are NOT part of a model (a
# sub-chain or a single process): the library always carries the whole column,
# the results of the missing processes are replaced by the caller's inputs and
# never reported, so these only have to be finite and pass the init checks.
# (Defaults follow static/metadata/parameters.yaml where it has one.)
_FILL_F64 = {
    "hru_area": 1.0, "hru_in_to_cf": 3630.0, "hru_lat": 40.0, "soil_moist_max": 1.0,
    "soil_rechr_max_frac": 1.0, "den_init": 0.1, "den_max": 0.6, "settle_const": 0.1,
    "carea_max": 0.6, "smidx_coef": 0.005, "smidx_exp": 0.3, "freeh2o_cap": 0.05,
    "emis_noppt": 0.757, "rad_trncf": 0.5, "snarea_thresh": 50.0, "potet_sublim": 0.5,
    "covden_sum": 0.5, "covden_win": 0.5, "jh_coef_hru": 13.0, "radj_sppt": 0.44,
    "radj_wppt": 0.5, "transp_tmax": 1.0, "dprst_depth_avg": 132.0, "va_open_exp": 0.001,
    "va_clos_exp": 0.001, "slowcoef_lin": 0.015, "slowcoef_sq": 0.1, "fastcoef_lin": 0.1,
    "fastcoef_sq": 0.8, "ssr2gw_exp": 1.0, "ssr2gw_rate": 0.1, "sat_threshold": 999.0,
    "gwflow_coef": 0.015, "radmax": 0.8, "jh_coef": 0.014, "dday_slope": 0.4, "dday_intcp": -40.0,
    "tmax_index": 50.0, "tmax_allsnow": 32.0, "tmax_allrain_offset": 1.0, "adjmix_rain": 1.0,
    "snow_cbh_adj": 1.0, "rain_cbh_adj": 1.0, "cecn_coef": 5.0, "radadj_intcp": 1.0,
    "ppt_rad_adj": 0.02, "mann_n": 0.04, "seg_depth": 1.0, "seg_length": 1000.0,
    "seg_slope": 0.001, "x_coef": 0.2, "albset_rna": 0.8, "albset_rnm": 0.6, "albset_sna": 0.05,
    "albset_snm": 0.2,
}
_FILL_I32 = {"cov_type": 1, "hru_type": 1, "hru_deplcrv": 1, "soil_type": 1, "transp_beg": 1,
             "transp_end": 13, "melt_force": 140, "melt_look": 90}


class _PinnedPool:
    """Page-locked host buffers for drained results.  Pinning memory costs about
    a second per GB (cudaHostAlloc), more than a 2-year CONUS-scale run takes, so
    buffers are handed back when an Engine closes and re-used by the next one
    (a model that is re-created per scenario, a service that keeps running).
    A buffer belongs to one Engine at a time."""

    CAP = 6 << 30  # bytes kept for re-use

    def __init__(self):
        self.free = {}
        self.held = 0

    def take(self, nbytes: int):
        lst = self.free.get(nbytes)
        if lst:
            self.held -= nbytes
            return lst.pop()
        import torch

        return torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)

    def give(self, buf):
        n = int(buf.numel())
        if self.held + n <= self.CAP:
            self.free.setdefault(n, []).append(buf)
            self.held += n


_PINNED = _PinnedPool()


class TiledBlock(dict):
    """Results of a run in the tiled layout ([day][tile][variable][tile_units],
    include/pws_b200.h: pws_output_tile): a dict that un-tiles an HRU variable
    into a [ndays, nhru] array the first time it is looked up.  Segment
    variables are ordinary entries."""

    def __init__(self, eng, of, oi):
        super().__init__()
        self._eng, self._of, self._oi = eng, of, oi
        self._f64 = {v: k for k, v in enumerate(eng.sel_f64)}
        self._i32 = {v: k for k, v in enumerate(eng.sel_i32)}

    def __missing__(self, var):
        if var in self._f64:
            a = np.ascontiguousarray(untile(self._of, self._f64[var], self._eng.nhru))
        elif var in self._i32:
            a = np.ascontiguousarray(untile(self._oi, self._i32[var], self._eng.nhru))
            if self._eng.reg.dtype(var) is np.bool_:
                a = a.astype(np.bool_)
        else:
            raise KeyError(var)
        self[var] = a
        return a

    def __contains__(self, var):
        return dict.__contains__(self, var) or var in self._f64 or var in self._i32

    def keys(self):
        return list(self._f64) + list(self._i32) + [k for k in dict.keys(self)
                                                     if k not in self._f64 and k not in self._i32]


class Engine:
    """A PRMS/NHM domain on one GPU.

    params: dict name -> ndarray with the reference's parameter names
    (static/metadata/parameters.yaml); `start`: first simulated day.
    """

    def __init__(self, params: dict, start, device: int = 0, dprst_flag: bool = True,
                 adjust_parameters: bool = True, channel: bool = True,
                 channel_chunk: int | None = None, fill_missing: bool = False,
                 forcing_period: int | None = None, soltab_device: bool | None = None,
                 flow_graph: tuple | None = None,
                 lanes: int | None = None):
        """`forcing_period` = P: a member-major parameter ensemble over a base
        domain of P HRUs (nhru = nmember * P); forcings are then [ndays, P] and
        shared by the members."""
        self.L = _lib.lib()
        self.reg = Registry.get()
        if fill_missing:
            params = self._filled(params)
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
        if soltab_device is not None:
            self._opt("soltab_device", 1.0 if soltab_device else 0.0)
        if lanes is not None:
            self._opt("lanes", lanes)
        self.ncol = self.nhru  # forcing columns
        if forcing_period:
            self._opt("forcing_period", int(forcing_period))
            self.ncol = int(forcing_period)
        self.tiled = False
        self.n_obs = 0
        if flow_graph is not None:
            # the segments are the nodes of a FlowGraph: (kind [nnodes], obs_col [nnodes], n_obs
            # [, flow_min [nnodes]])
            kind, col, self.n_obs = flow_graph[:3]
            kind = np.ascontiguousarray(kind, dtype=np.int32)
            col = np.ascontiguousarray(col, dtype=np.int32)
            if kind.shape != (self.nseg,) or col.shape != (self.nseg,):
                raise ValueError("flow_graph: kind and obs_col must be [nnodes]")
            self._check(self.L.pws_set_flow_graph(self.h, kind.ctypes.data, col.ctypes.data, int(self.n_obs)))
            if len(flow_graph) > 3 and flow_graph[3] is not None:
                fmin = np.ascontiguousarray(flow_graph[3], dtype=np.float64)
                if fmin.shape != (self.nseg,):
                    raise ValueError("flow_graph: flow_min must be [nnodes]")
                self._check(self.L.pws_set_flow_node_params(self.h, fmin.ctypes.data))
        self._check(self.L.pws_init(self.h))
        self.day = 0
        self.sel_hru: list[str] = []
        self.sel_seg: list[str] = []

    @staticmethod
    def _filled(params: dict) -> dict:
        """`params` completed with stand-in values (see _FILL_F64)."""
        p = dict(params)
        nhru = None
        for k in ("hru_area", "hru_segment", "nhm_id", "hru_type", "cov_type"):
            if k in p:
                nhru = int(np.asarray(p[k]).shape[-1])
                break
        if nhru is None:
            raise ValueError("cannot tell nhru: no [nhru] parameter was given")
        for k in HRU_PARAMS_F64:
            p.setdefault(k, np.full(nhru, _FILL_F64.get(k, 0.0)))
        for k in MON_PARAMS_F64:
            p.setdefault(k, np.full((12, nhru), _FILL_F64.get(k, 0.0)))
        for k in HRU_PARAMS_I32:
            p.setdefault(k, np.full(nhru, _FILL_I32.get(k, 0), dtype=np.int64))
        for k in MON_PARAMS_I32:
            p.setdefault(k, np.zeros((12, nhru), dtype=np.int64))
        for k in SCALAR_OPTIONS:
            p.setdefault(k, np.array([_FILL_F64[k]]))
        p.setdefault("snarea_curve", np.linspace(0.05, 1.0, 11))
        if "tosegment" in p:
            nseg = int(np.asarray(p["tosegment"]).shape[0])
            for k in SEG_PARAMS_F64:
                p.setdefault(k, np.full(nseg, _FILL_F64.get(k, 0.0)))
            p.setdefault("segment_type", np.zeros(nseg, dtype=np.int64))
        return p

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
        """(column kernel launches, channel kernel launches) of the last run."""
        a, b = C.c_int(), C.c_int()
        self.L.pws_last_launches(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def last_blocks(self):
        """(blocks of days, column launches per block) of the last run."""
        a, b = C.c_int(), C.c_int()
        self.L.pws_last_blocks(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def release_buffers(self):
        """Hand the page-locked result buffers back to the pool (the arrays
        returned by earlier run_host(pinned_outputs=True) calls must not be used
        any more)."""
        for buf in self.__dict__.pop("_pinned", []):
            _PINNED.give(buf)
        self.__dict__.pop("_out_caches", None)

    def close(self):
        if getattr(self, "h", None):
            self.L.pws_destroy(self.h)
            self.h = None
            self.release_buffers()

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

    def select_accumulations(self, hru_vars=(), seg_vars=()):
        """Budget terms summed per unit on the device over every day computed
        from now on (base/budget.py:262-269); they need not be selected outputs."""
        hru_vars, seg_vars = list(hru_vars), list(seg_vars) if self.nseg else []
        hi = np.array([self.reg.hru_index[v] for v in hru_vars], dtype=np.int32)
        si = np.array([self.reg.seg_index[v] for v in seg_vars], dtype=np.int32)
        self._check(self.L.pws_select_accumulations(self.h, len(hi), hi.ctypes.data, len(si),
                                                    si.ctypes.data))
        self.acc_hru, self.acc_seg = hru_vars, seg_vars

    def accumulations(self) -> dict:
        """var -> per-unit sum over the days computed since the last reset."""
        hv, sv = getattr(self, "acc_hru", []), getattr(self, "acc_seg", [])
        ah = np.empty((len(hv), self.nhru))
        as_ = np.empty((len(sv), self.nseg))
        self._check(self.L.pws_get_accumulations(self.h, ah.ctypes.data if hv else None,
                                                 as_.ctypes.data if sv else None))
        out = {v: ah[k] for k, v in enumerate(hv)}
        out.update({v: as_[k] for k, v in enumerate(sv)})
        return out

    def set_accumulate(self, on: bool):
        self._opt("accumulate", 1.0 if on else 0.0)

    def soltab(self):
        out = {}
        for k, name in enumerate(("soltab_potsw", "soltab_horad_potsw", "soltab_sunhrs")):
            a = np.empty((366, self.nhru))
            self._check(self.L.pws_get_soltab(self.h, k, a.ctypes.data))
            out[name] = a
        return out

    # ---- runs
    def run_host(self, prcp, tmax, tmin, budgets: bool = True, pinned_outputs: bool = False,
                 reuse_buffers: bool = False):
        """Advance len(prcp) days with forcings / outputs in host memory
        (time-blocked, H2D / compute / D2H overlapped).  Returns
        (dict var -> [ndays, n] ndarray with the reference's dtype,
        budget violation counts [ndays, 6] or None).  Pageable arrays are
        staged through pinned buffers inside the library; with
        `pinned_outputs` the result arrays are allocated page-locked (through
        torch) so the drain needs no staging copy; with `reuse_buffers` the
        result arrays of the previous call with the same shapes are written
        again (the caller must be done with them), which avoids first-touch
        page faults on every block of a long run; a number > 1 names another
        set of buffers, so a caller can alternate between two.

        Forcings that are all float32 (the storage type of the reference's
        forcing netCDF files) are uploaded as they are -- half the PCIe
        traffic -- and widened on the device (pws_run_host_f32): the same
        bits as passing `.astype(float64)`."""
        f32 = all(getattr(a, "dtype", None) == np.float32 for a in (prcp, tmax, tmin))
        ft = np.float32 if f32 else np.float64
        prcp = np.ascontiguousarray(prcp, dtype=ft)
        tmax = np.ascontiguousarray(tmax, dtype=ft)
        tmin = np.ascontiguousarray(tmin, dtype=ft)
        nd = prcp.shape[0]
        if not (prcp.shape == tmax.shape == tmin.shape == (nd, self.ncol)):
            raise ValueError(f"forcings must be [ndays, {self.ncol}]")
        cal = calendar(self.start + self.day, nd)
        if self.tiled:
            # [day][tile][variable][tile_units], the last tile padded (pws_output_tile)
            H, padded = self.output_tile()
            shapes = ((nd, padded // H, len(self.sel_f64), H), (nd, padded // H, len(self.sel_i32), H),
                      (nd, len(self.sel_seg), self.nseg))
        else:
            shapes = ((nd, len(self.sel_f64), self.nhru), (nd, len(self.sel_i32), self.nhru),
                      (nd, len(self.sel_seg), self.nseg))
        key = (shapes, bool(pinned_outputs))
        caches = self.__dict__.setdefault("_out_caches", {})
        cache = caches.get(int(reuse_buffers)) if reuse_buffers else None
        if cache is not None and cache[0] == key:
            of, oi, os_ = cache[1]
        else:
            if pinned_outputs:
                def alloc(shape, dt):
                    n = int(np.prod(shape)) * np.dtype(dt).itemsize
                    buf = _PINNED.take(n)
                    self.__dict__.setdefault("_pinned", []).append(buf)
                    return buf.numpy()[:n].view(dt).reshape(shape)
                of = alloc(shapes[0], np.float64)
                oi = alloc(shapes[1], np.int32)
                os_ = alloc(shapes[2], np.float64)
            else:
                of = np.empty(shapes[0])
                oi = np.empty(shapes[1], dtype=np.int32)
                os_ = np.empty(shapes[2])
            if reuse_buffers:
                caches[int(reuse_buffers)] = (key, (of, oi, os_))
        viol = np.zeros((nd, 6), dtype=np.int32) if budgets else None
        run = self.L.pws_run_host_f32 if f32 else self.L.pws_run_host
        self._check(run(
            self.h, nd, cal.ctypes.data, prcp.ctypes.data, tmax.ctypes.data, tmin.ctypes.data,
            of.ctypes.data, oi.ctypes.data, os_.ctypes.data,
            viol.ctypes.data if budgets else None))
        self.day += nd
        if self.tiled:
            res = TiledBlock(self, of, oi)
        else:
            res = {v: of[:, k, :] for k, v in enumerate(self.sel_f64)}
            for k, v in enumerate(self.sel_i32):
                a = oi[:, k, :]
                res[v] = a.astype(np.bool_) if self.reg.dtype(v) is np.bool_ else a
        res.update({v: os_[:, k, :] for k, v in enumerate(self.sel_seg)})
        return res, viol

    # bits of `process_mask` (= the library's budget numbering)
    PROCESS_BITS = {"PRMSCanopy": 0, "PRMSSnow": 1, "PRMSRunoff": 2, "PRMSSoilzone": 3,
                    "PRMSGroundwater": 4, "PRMSChannel": 5}

    def run_host_inputs(self, inputs: dict, process_names, budgets: bool = True,
                        ndays: int | None = None):
        """Advance a SUBSET of the chain (`process_names`), every variable the
        missing processes would have produced supplied in `inputs`
        (name -> [ndays, nhru]; names = the reference's get_inputs()).  Returns
        like run_host."""
        names = [k for k in inputs if k in self.reg.input_index]
        unknown = [k for k in inputs if k not in self.reg.input_index]
        if unknown:
            raise ValueError(f"unknown inputs {unknown}")
        arrs = [np.ascontiguousarray(inputs[k], dtype=np.float64) for k in names]
        nd = arrs[0].shape[0] if arrs else int(ndays or 0)
        for k, a in zip(names, arrs):
            if a.shape != (nd, self.nhru):
                raise ValueError(f"input '{k}' must be [ndays, nhru] = {(nd, self.nhru)}, got {a.shape}")
        mask = 0
        for p in process_names:
            if p in self.PROCESS_BITS:
                mask |= 1 << self.PROCESS_BITS[p]
        cal = calendar(self.start + self.day, nd)
        idx = np.array([self.reg.input_index[k] for k in names], dtype=np.int32)
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        of = np.empty((nd, len(self.sel_f64), self.nhru))
        oi = np.empty((nd, len(self.sel_i32), self.nhru), dtype=np.int32)
        os_ = np.empty((nd, len(self.sel_seg), self.nseg))
        viol = np.zeros((nd, 6), dtype=np.int32) if budgets else None
        self._check(self.L.pws_run_host_inputs(
            self.h, nd, cal.ctypes.data, len(arrs), idx.ctypes.data, C.cast(ptrs, C.c_void_p), mask,
            of.ctypes.data, oi.ctypes.data, os_.ctypes.data,
            viol.ctypes.data if budgets else None))
        self.day += nd
        res = {v: of[:, k, :] for k, v in enumerate(self.sel_f64)}
        for k, v in enumerate(self.sel_i32):
            a = oi[:, k, :]
            res[v] = a.astype(np.bool_) if self.reg.dtype(v) is np.bool_ else a
        if mask & (1 << self.PROCESS_BITS["PRMSChannel"]):
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

    # ---- tiled result layout of run_device (include/pws_b200.h: pws_output_tile)
    def set_tiled_outputs(self, on: bool = True):
        """Results of run_device as [day][tile][variable][tile_units] (every HRU
        variable must be selected in registry order)."""
        self._opt("tiled_outputs", 1.0 if on else 0.0)
        self.tiled = bool(on)

    def output_tile(self):
        """(tile_units, padded_units): units per tile and per padded variable row."""
        t, p = C.c_int(), C.c_int64()
        self._check(self.L.pws_output_tile(self.h, C.byref(t), C.byref(p)))
        return t.value, p.value

    def tiled_view(self, out, var: str):
        """[nd, nhru] strided view of variable `var` in a tiled result tensor
        `out` = [nd, tiles, n_vars_of_its_kind, tile_units] (torch or numpy)."""
        sel = self.sel_f64 if self.reg.dtype(var) is np.float64 else self.sel_i32
        return untile(out, sel.index(var), self.nhru)

    FLOW_GRAPH_OUT = ("outflows", "node_upstream_inflows", "node_outflows", "node_storage_changes",
                      "node_sink_source")

    def run_flow_graph(self, inflows, obs=None) -> dict:
        """FlowGraph.calculate for len(inflows) days (pws_run_flow_graph_host):
        inflows [ndays, nnodes], obs [ndays, n_obs].  Returns var -> [ndays, nnodes]."""
        lat = np.ascontiguousarray(inflows, dtype=np.float64)
        nd = lat.shape[0]
        if lat.shape != (nd, self.nseg):
            raise ValueError(f"inflows must be [ndays, {self.nseg}]")
        if self.n_obs:
            obs = np.ascontiguousarray(obs, dtype=np.float64)
            if obs.shape != (nd, self.n_obs):
                raise ValueError(f"obs must be [ndays, {self.n_obs}]")
        out = np.empty((nd, len(self.FLOW_GRAPH_OUT), self.nseg))
        self._check(self.L.pws_run_flow_graph_host(
            self.h, nd, lat.ctypes.data, obs.ctypes.data if self.n_obs else None, out.ctypes.data))
        self.day += nd
        return {v: out[:, k, :] for k, v in enumerate(self.FLOW_GRAPH_OUT)}

    def run_channel_device(self, nd, d_seg_lat, d_seg_out, sync=True):
        self._check(self.L.pws_run_channel_device(self.h, nd, d_seg_lat, d_seg_out))
        if sync:
            self.sync()

    def sync(self):
        self._check(self.L.pws_sync(self.h))

    # ---- state / instrumentation
    @property
    def seg_states(self):
        L = self.L
        return [L.pws_seg_state_name(i).decode() for i in range(L.pws_n_seg_state())] if self.nseg else []

    def get_state(self, name):
        a = np.empty(self.nseg if name in self.seg_states else self.nhru)
        self._check(self.L.pws_get_state(self.h, name.encode(), a.ctypes.data))
        return a

    def set_state(self, name, val):
        a = np.ascontiguousarray(val, dtype=np.float64)
        assert a.shape == ((self.nseg if name in self.seg_states else self.nhru),)
        self._check(self.L.pws_set_state(self.h, name.encode(), a.ctypes.data))

    def save_state(self) -> dict:
        """Every prognostic value of the domain (the union of the processes'
        get_restart_variables(), prms_snow.py:269-293, plus PRMSChannel's
        outflow_ts / seg_inflow) and the day it belongs to: a restart point."""
        st = {name: self.get_state(name) for name in self.reg.states + self.seg_states}
        st["_day"] = np.array(self.day)
        return st

    def load_state(self, st: dict):
        for name in self.reg.states + self.seg_states:
            self.set_state(name, st[name])
        self.day = int(st["_day"])

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


class EtEngine:
    """PRMSEt on the GPU (pws_et_host): stateless, driven in blocks of days like
    an Engine so the Model treats it the same way."""

    INPUTS = ("potet", "hru_impervevap", "hru_intcpevap", "snow_evap", "dprst_evap_hru", "perv_actet")
    VARS = ("unused_potet", "hru_perv_actet", "hru_actet")

    def __init__(self, params: dict, start, device: int = 0):
        self.L = _lib.lib()
        self.reg = Registry.get()
        self.device = device
        self.imperv = np.ascontiguousarray(params["hru_percent_imperv"], dtype=np.float64)
        self.dprst = np.ascontiguousarray(params["dprst_frac"], dtype=np.float64)
        self.nhru, self.nseg, self.day = int(self.imperv.shape[0]), 0, 0
        self.start = np.datetime64(start, "D")
        self.sel_seg = []

    def run(self, inputs: dict):
        arrs = [np.ascontiguousarray(inputs[k], dtype=np.float64) for k in self.INPUTS]
        nd = arrs[0].shape[0]
        for k, a in zip(self.INPUTS, arrs):
            if a.shape != (nd, self.nhru):
                raise ValueError(f"input '{k}' must be [ndays, nhru] = {(nd, self.nhru)}, got {a.shape}")
        outs = [np.empty((nd, self.nhru)) for _ in self.VARS]
        rc = self.L.pws_et_host(self.device, nd, self.nhru, *[a.ctypes.data for a in arrs],
                                self.imperv.ctypes.data, self.dprst.ctypes.data,
                                *[o.ctypes.data for o in outs])
        if rc != 0:
            raise _lib.PwsLibraryError(self.L.pws_last_error(None).decode())
        self.day += nd
        return dict(zip(self.VARS, outs))

    def close(self):
        pass
