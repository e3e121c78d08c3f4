"""The eight PRMS/NHM process classes with the reference's constructor
signatures, static descriptors and public variable names
(pywatershed/base/process.py:19-592, conservative_process.py:13-210 and the
per-process files cited on each class).

On the reference a process computes itself in `calculate()`; here the whole
vertical column runs fused on the GPU, so a process object is a *view*: its
public variables are stable numpy arrays (same names, dtypes and shapes as the
reference) that the owning Model refreshes for the current day, plus a Budget.
`calc_method` accepts "cuda" (default); "numpy"/"numba" name CPU paths this
package does not have and raise.
"""
from __future__ import annotations

import pathlib as pl
from warnings import warn

import numpy as np

from ._process_meta import PROCESS_META, VARIABLE_META
from .budget import Budget

_DTYPES = {"float64": np.float64, "int32": np.int32, "bool": np.bool_, "int64": np.int64}


def _full(shape, fill, dtype):
    """np.full; large arrays (GBs at CONUS size) are first touched by a few threads."""
    a = np.empty(shape, dtype=dtype)
    if a.nbytes < (64 << 20):
        a[...] = fill
        return a
    from concurrent.futures import ThreadPoolExecutor

    nthr = 8
    step = -(-shape[0] // nthr)

    def fill_rows(k):
        a[k * step:(k + 1) * step] = fill

    with ThreadPoolExecutor(max_workers=nthr) as ex:
        list(ex.map(fill_rows, range(nthr)))
    return a


class TimeseriesArray:
    """All-time variable with a `.current` row (base/timeseries.py:8-70).

    `.data` is [ntime or 366, nhru].  The reference computes these arrays up
    front (SolarGeometry at construction, Atmosphere on the first advance); here
    they are materialised when somebody looks: the array is allocated on first
    access and `loader(self)` -- the owning Model -- fills whatever rows were
    computed on the device without being brought to the host (a fast
    `Model.run()` does not drain 18 [ntime, nhru] arrays nobody reads)."""

    def __init__(self, control, var_name, array=None, doy_indexed=False, shape=None,
                 dtype=np.float64, fill=np.nan):
        self.name = "TimeseriresArray"
        self.control = control
        self.variable_name = var_name
        self._array = array
        self._shape = tuple(array.shape) if array is not None else tuple(shape)
        self._dtype = array.dtype if array is not None else np.dtype(dtype)
        self._fill = fill
        self._doy = doy_indexed
        self._current = np.full(self._shape[1], fill, dtype=self._dtype)
        self._loader = None
        self._stale = False   # rows exist on the device side that `_array` does not hold

    @property
    def shape(self):
        return self._shape

    @property
    def allocated(self) -> bool:
        return self._array is not None

    @property
    def data(self):
        if self._array is None:
            self._array = _full(self._shape, self._fill, self._dtype)
        if self._stale and self._loader is not None:
            self._stale = False
            self._loader(self)
        return self._array

    @data.setter
    def data(self, value):
        self._array = value
        self._stale = False

    def set_rows(self, t0, rows):
        """Rows [t0, t0 + len(rows)) as delivered by a block; kept only when the
        array exists already (otherwise they are recomputed on demand)."""
        if self._array is None:
            self._stale = True
        else:
            self._array[t0:t0 + len(rows)] = rows

    def mark_stale(self):
        self._stale = True

    def set_current(self, row):
        self._current[:] = row

    def advance(self):
        if self._doy:
            ind = self.control.current_doy - 1
        else:
            ind = max(self.control.itime_step, 0)
        self._current[:] = self.data[ind, :]

    @property
    def current(self):
        return self._current


class Process:
    """Base class: self-describing, variables allocated from metadata."""

    _all_time = False  # SolarGeometry / Atmosphere keep [ntime, nhru] arrays

    def __init__(self, control, discretization, parameters, **kwargs):
        self.name = type(self).__name__
        self.control = control
        self.discretization = discretization
        self._params_in = parameters
        self._model = None
        meta = PROCESS_META[self.name]
        from .parameters import Parameters, as_param_dict

        merged = {}
        for p in (discretization, parameters):
            if p is not None:
                merged.update(as_param_dict(p))
        # declared by the reference but never read on this path (gwstor_min,
        # tosegment_nhm, obsin_segment, ... ) are not required
        from . import engine as _e

        used = set(_e.HRU_PARAMS_F64 + _e.MON_PARAMS_F64 + _e.HRU_PARAMS_I32 + _e.MON_PARAMS_I32
                   + _e.SEG_PARAMS_F64 + _e.SEG_PARAMS_I32 + _e.SCALAR_OPTIONS
                   + ("snarea_curve", "temp_units"))
        missing = [k for k in meta["parameters"] if k not in merged and k in used]
        if missing:
            # base/process.py:233-237
            raise ValueError(f"{self.name}: parameters missing: {sorted(missing)}")
        self._param_dict = merged
        self.params = parameters
        self.nhru = 0
        for key in ("hru_area", "hru_percent_imperv", "nhm_id"):
            if key in merged:
                self.nhru = int(np.asarray(merged[key]).shape[0])
                break
        self.nsegment = int(np.asarray(merged["tosegment"]).shape[0]) if "tosegment" in merged else 0
        self._input_variables_dict = {k: kwargs.get(k) for k in meta["inputs"]}
        self._vars = {}
        self._netcdf_initialized = False
        self._verbose = kwargs.get("verbose")
        self._set_options(kwargs)
        self._allocate()

    # ---- descriptors (static in the reference, base/process.py:144-187)
    @classmethod
    def get_dimensions(cls) -> tuple:
        return PROCESS_META[cls.__name__]["dimensions"]

    @classmethod
    def get_parameters(cls) -> tuple:
        return PROCESS_META[cls.__name__]["parameters"]

    @classmethod
    def get_inputs(cls) -> tuple:
        return PROCESS_META[cls.__name__]["inputs"]

    @classmethod
    def get_variables(cls) -> tuple:
        return PROCESS_META[cls.__name__]["variables"]

    @classmethod
    def get_init_values(cls) -> dict:
        return {k: (np.nan if v is None else v)
                for k, v in PROCESS_META[cls.__name__]["init_values"].items()}

    @classmethod
    def get_restart_variables(cls) -> tuple:
        return PROCESS_META[cls.__name__].get("restart_variables", ())

    @classmethod
    def description(cls) -> dict:
        return {"class_name": cls.__name__, "mass_budget_terms": getattr(
                    cls, "get_mass_budget_terms", lambda: None)(),
                "inputs": cls.get_inputs(), "variables": cls.get_variables(),
                "parameters": cls.get_parameters()}

    @property
    def dimensions(self):
        return self.get_dimensions()

    @property
    def parameters(self):
        return self.get_parameters()

    @property
    def inputs(self):
        return self.get_inputs()

    @property
    def variables(self):
        return self.get_variables()

    @property
    def init_values(self):
        return self.get_init_values()

    # ---- options: per-process kwargs override control (base/process.py:339-360)
    def _set_options(self, kw):
        # base/process.py:339-358: an option exists on a process only when its
        # constructor names it; the constructor argument wins over the control
        import inspect

        opts = getattr(self.control, "options", {}) or {}
        own = inspect.signature(type(self).__init__).parameters
        for name in ("budget_type", "calc_method", "dprst_flag", "adjust_parameters"):
            val = kw.get(name)
            if name not in own and val is None:
                setattr(self, "_" + name, None)
                continue
            if val is None and name in opts:
                val = opts[name]
            setattr(self, "_" + name, val)
        cm = self._calc_method
        if cm not in (None, "cuda"):
            if cm not in ("numpy", "numba", "fortran"):
                raise ValueError(f"{self.name}: invalid calc_method '{cm}'")
            if kw.get("calc_method") is not None:
                # asked for by name in the constructor: there is no such path here
                raise ValueError(
                    f"{self.name}: calc_method='{cm}' names a CPU path of the reference; "
                    "pywatershed_b200 only has calc_method='cuda' (no CPU fallback)")
            # inherited from a control file written for the reference: say what happens instead
            warn(
                f"{self.name}: control option calc_method='{cm}' names a CPU path of the reference; "
                "pywatershed_b200 computes with calc_method='cuda'", UserWarning, stacklevel=3)
        self._calc_method = "cuda"

    def _n_units(self, var):
        dims = VARIABLE_META[var]["dims"]
        return self.nsegment if "nsegment" in dims else self.nhru

    def _allocate(self):
        init = self.get_init_values()
        for var in self.get_variables():
            dt = _DTYPES[VARIABLE_META[var]["type"]]
            n = self._n_units(var)
            fill = init[var]
            if dt is not np.float64 and (fill is None or (isinstance(fill, float) and np.isnan(fill))):
                fill = 0
            self._vars[var] = np.full(n, fill, dtype=dt)

    # ---- access (base/accessor.py:1-15)
    def __getitem__(self, key):
        if key in self._vars:
            return self._vars[key]
        if key in self._param_dict:
            return self._param_dict[key]
        return getattr(self, key)

    def __getattr__(self, key):
        d = self.__dict__
        if "_vars" in d and key in d["_vars"]:
            return d["_vars"][key]
        if "_param_dict" in d and key in d["_param_dict"]:
            return d["_param_dict"][key]
        if "_input_variables_dict" in d and key in d["_input_variables_dict"]:
            return self._input_current(key)
        raise AttributeError(key)

    def _input_current(self, key):
        """An input is the producing process's array, shared by reference
        (base/process.py:362-377): the current-day [n] values."""
        model = self.__dict__.get("_model")
        if model is not None:
            for proc in model.processes.values():
                if proc is not self and key in proc._vars:
                    v = proc._vars[key]
                    return v.current if isinstance(v, TimeseriesArray) else v
        raise AttributeError(f"{self.name}: input '{key}' is only available under a Model")

    def set_input_to_adapter(self, input_variable_name: str, adapter):
        self._input_variables_dict[input_variable_name] = adapter

    # ---- stepping a process on its own (base/process.py:379-422): the caller
    # advances the Control, then proc.advance(); proc.calculate(dt); proc.output().
    # Behind it sits a private one-process Model whose missing upstream
    # processes are the process's input arrays / files.
    def _need_model(self):
        if self._model is None:
            from .model import Model

            self._model = Model._around(self)
            self._owns_model = True
        return self._model

    def advance(self):
        m = self._need_model()
        if getattr(self, "_owns_model", False):
            m.advance()

    def calculate(self, time_length: float = 1.0, **kwargs):
        m = self._need_model()
        if getattr(self, "_owns_model", False):
            m.calculate()

    def output(self):
        m = self._need_model()
        if getattr(self, "_owns_model", False):
            m.output()

    def finalize(self):
        pass

    def initialize_netcdf(self, *args, **kwargs):
        self._need_model()
        if self._netcdf_initialized:
            warn(f"{self.name} class previously initialized netcdf output", RuntimeWarning)
            return
        self._model._initialize_netcdf_for(self, *args, **kwargs)
        self._netcdf_initialized = True


class ConservativeProcess(Process):
    _budget_basis = "unit"

    def __init__(self, control, discretization, parameters, **kwargs):
        super().__init__(control, discretization, parameters, **kwargs)
        terms = self.get_mass_budget_terms()
        self.budget = None
        if self._budget_type is not None:
            self.budget = Budget(control, self, terms, self._budget_type, basis=self._budget_basis,
                                 description=self.name)

    @classmethod
    def get_mass_budget_terms(cls) -> dict:
        return PROCESS_META[cls.__name__]["mass_budget_terms"]


# --------------------------------------------------------------------------
class PRMSSolarGeometry(Process):
    """atmosphere/prms_solar_geometry.py:60-68"""
    _all_time = True

    def __init__(self, control, discretization, parameters, verbose: bool = None,
                 from_prms_file=None, from_nc_files_dir=None):
        if from_prms_file is not None or from_nc_files_dir is not None:
            raise NotImplementedError("PRMSSolarGeometry from files is not implemented: the "
                                      "tables are computed (pws_init)")
        super().__init__(control, discretization, parameters, verbose=verbose)

    def _allocate(self):
        for var in self.get_variables():
            self._vars[var] = TimeseriesArray(self.control, var, shape=(366, self.nhru),
                                              doy_indexed=True)


class PRMSAtmosphere(Process):
    """atmosphere/prms_atmosphere.py:88-99"""
    _all_time = True

    def __init__(self, control, discretization, parameters, prcp=None, tmax=None, tmin=None,
                 soltab_potsw=None, soltab_horad_potsw=None, verbose: bool = None):
        super().__init__(control, discretization, parameters, prcp=prcp, tmax=tmax, tmin=tmin,
                         soltab_potsw=soltab_potsw, soltab_horad_potsw=soltab_horad_potsw,
                         verbose=verbose)

    def _allocate(self):
        nt = self.control.n_times
        init = self.get_init_values()
        for var in self.get_variables():
            dt = _DTYPES[VARIABLE_META[var]["type"]]
            fill = init[var]
            if dt is not np.float64 and isinstance(fill, float) and np.isnan(fill):
                fill = 0
            self._vars[var] = TimeseriesArray(self.control, var, shape=(nt, self.nhru), dtype=dt,
                                              fill=fill)


class PRMSCanopy(ConservativeProcess):
    """hydrology/prms_canopy.py:67-83"""

    def __init__(self, control, discretization, parameters, pk_ice_prev=None, freeh2o_prev=None,
                 transp_on=None, hru_ppt=None, hru_rain=None, hru_snow=None, potet=None,
                 pptmix=None, budget_type=None, calc_method=None, verbose: bool = None):
        super().__init__(control, discretization, parameters, **_kw(locals()))


class PRMSSnow(ConservativeProcess):
    """hydrology/prms_snow.py:122-145"""

    def __init__(self, control, discretization, parameters, orad_hru=None, soltab_horad_potsw=None,
                 swrad=None, hru_intcpevap=None, hru_ppt=None, potet=None, pptmix=None, prmx=None,
                 tavgc=None, tmaxc=None, tminc=None, net_ppt=None, net_rain=None, net_snow=None,
                 transp_on=None, budget_type=None, calc_method=None, verbose: bool = None):
        super().__init__(control, discretization, parameters, **_kw(locals()))


class PRMSRunoff(ConservativeProcess):
    """hydrology/prms_runoff.py:77-100"""

    def __init__(self, control, discretization, parameters, soil_lower_prev=None,
                 soil_rechr_prev=None, net_ppt=None, net_rain=None, net_snow=None, potet=None,
                 snowmelt=None, snow_evap=None, pkwater_equiv=None, pptmix_nopack=None,
                 snowcov_area=None, through_rain=None, hru_intcpevap=None, intcp_changeover=None,
                 dprst_flag: bool = None, budget_type=None, calc_method=None,
                 verbose: bool = None):
        super().__init__(control, discretization, parameters, **_kw(locals()))


class PRMSSoilzone(ConservativeProcess):
    """hydrology/prms_soilzone.py:81-102"""

    def __init__(self, control, discretization, parameters, dprst_evap_hru=None,
                 dprst_seep_hru=None, hru_impervevap=None, hru_intcpevap=None, infil_hru=None,
                 sroff=None, sroff_vol=None, potet=None, transp_on=None, snow_evap=None,
                 snowcov_area=None, dprst_flag: bool = None, budget_type=None, calc_method=None,
                 adjust_parameters="warn", verbose: bool = None):
        if adjust_parameters not in ("warn", "error", "no"):
            raise ValueError("adjust_parameters must be one of 'warn', 'error', 'no'")
        super().__init__(control, discretization, parameters, **_kw(locals()))


class PRMSGroundwater(ConservativeProcess):
    """hydrology/prms_groundwater.py:52-64"""

    def __init__(self, control, discretization, parameters, soil_to_gw=None, ssr_to_gw=None,
                 dprst_seep_hru=None, dprst_flag: bool = None, budget_type=None,
                 calc_method=None, verbose: bool = None):
        super().__init__(control, discretization, parameters, **_kw(locals()))


class PRMSChannel(ConservativeProcess):
    """hydrology/prms_channel.py:92-104; global-basis budget (:115)"""
    _budget_basis = "global"

    def __init__(self, control, discretization, parameters, sroff_vol=None, ssres_flow_vol=None,
                 gwres_flow_vol=None, budget_type=None, calc_method=None,
                 adjust_parameters="warn", verbose: bool = None):
        super().__init__(control, discretization, parameters, **_kw(locals()))


class PRMSEt(Process):
    """hydrology/prms_et.py:22-48.  The reference builds this process's Budget
    with keyword arguments its Budget class no longer has (:57-76), so only
    budget_type=None / "defer" can be constructed there; here any budget_type is
    accepted and no budget is kept."""

    def __init__(self, control, discretization, parameters, potet=None, hru_impervevap=None,
                 hru_intcpevap=None, snow_evap=None, dprst_evap_hru=None, perv_actet=None,
                 verbose: bool = False, budget_type="defer"):
        super().__init__(control, discretization, parameters, potet=potet,
                         hru_impervevap=hru_impervevap, hru_intcpevap=hru_intcpevap,
                         snow_evap=snow_evap, dprst_evap_hru=dprst_evap_hru, perv_actet=perv_actet,
                         verbose=verbose)
        self.budget = None


class PRMSRunoffNoDprst(PRMSRunoff):
    """hydrology/prms_runoff_no_dprst.py:68-91: PRMSRunoff without depression storage"""

    def __init__(self, control, discretization, parameters, soil_lower_prev=None,
                 soil_rechr_prev=None, net_ppt=None, net_rain=None, net_snow=None, potet=None,
                 snowmelt=None, snow_evap=None, pkwater_equiv=None, pptmix_nopack=None,
                 snowcov_area=None, through_rain=None, hru_intcpevap=None, intcp_changeover=None,
                 budget_type=None, calc_method=None, verbose: bool = None):
        kw = _kw(locals())
        ConservativeProcess.__init__(self, control, discretization, parameters, dprst_flag=False, **kw)


class PRMSSoilzoneNoDprst(PRMSSoilzone):
    """hydrology/prms_soilzone_no_dprst.py:62-79"""

    def __init__(self, control, discretization, parameters, hru_impervevap=None,
                 hru_intcpevap=None, infil_hru=None, sroff=None, sroff_vol=None, potet=None,
                 transp_on=None, snow_evap=None, snowcov_area=None, budget_type=None,
                 calc_method=None, adjust_parameters="warn", verbose: bool = None):
        if adjust_parameters not in ("warn", "error", "no"):
            raise ValueError("adjust_parameters must be one of 'warn', 'error', 'no'")
        kw = _kw(locals())
        ConservativeProcess.__init__(self, control, discretization, parameters, dprst_flag=False, **kw)


class PRMSGroundwaterNoDprst(PRMSGroundwater):
    """hydrology/prms_groundwater_no_dprst.py:42-51"""

    def __init__(self, control, discretization, parameters, soil_to_gw=None, ssr_to_gw=None,
                 budget_type=None, calc_method=None, verbose: bool = None):
        kw = _kw(locals())
        ConservativeProcess.__init__(self, control, discretization, parameters, dprst_flag=False, **kw)


def check_adjusted_parameters(pd: dict, soilzone_mode, channel_mode, dprst_flag=True):
    """The parameter checks PRMSSoilzone and PRMSChannel make at construction
    (prms_soilzone.py:361-463, prms_channel.py:246-260), on the parameters, in
    the reference's order and with its messages: mode "warn" -> UserWarning (the
    device then applies the same edits), "error" -> the warnings, then
    ValueError, "no" / None -> nothing.  The edits cascade, so each mask is
    evaluated on the values the previous edits left."""
    def arr(k):
        return np.asarray(pd[k], dtype=np.float64)

    if soilzone_mode in ("warn", "error") and "soil_moist_max" in pd:
        hru_type = np.asarray(pd["hru_type"])
        soil_moist_max = arr("soil_moist_max").copy()
        soil_rechr_max = arr("soil_rechr_max_frac") * soil_moist_max
        soil_moist = arr("soil_moist_init_frac") * soil_moist_max
        soil_rechr = arr("soil_rechr_init_frac") * soil_rechr_max
        sat = arr("sat_threshold").copy()
        sat[(hru_type == 0) | (hru_type == 2)] = 0.0
        ssres_stor = arr("ssstor_init_frac") * sat
        ssres_stor[(hru_type == 0) | (hru_type == 2)] = 0.0
        edited = False

        def check(mask, msg):
            nonlocal edited
            if mask.any():
                edited = True
                if msg:
                    warn(f"{msg} at indices: {np.where(mask)[0]}", UserWarning, stacklevel=4)
            return mask

        m = check(soil_moist_max < 1.0e-5, "soil_moist_max < 1.0e-5, set to 1.0e-5")
        soil_moist_max = np.where(m, 1.0e-5, soil_moist_max)
        m = check(soil_rechr_max < 1.0e-5, "soil_rechr_max < 1.0e-5, set to 1.0e-5")
        soil_rechr_max = np.where(m, 1.0e-5, soil_rechr_max)
        m = check(soil_rechr_max > soil_moist_max,
                  "soil_rechr_max > soil_moist_max, soil_rechr_max set to soil_moist_max")
        soil_rechr_max = np.where(m, soil_moist_max, soil_rechr_max)
        m = check(soil_rechr > soil_rechr_max,
                  "soil_rechr_init > soil_rechr_max, setting soil_rechr_init to soil_rechr_max")
        soil_rechr = np.where(m, soil_rechr_max, soil_rechr)
        m = check(soil_moist > soil_moist_max,
                  "soil_moist_init > soil_moist_max, setting soil_moist to soil_moist max")
        soil_moist = np.where(m, soil_moist_max, soil_moist)
        m = check(soil_rechr > soil_moist, "soil_rechr > soil_moist, setting soil_rechr to soil_moist")
        # (the ssres_stor edit counts for "error" but is not announced, prms_soilzone.py:440-452)
        check(ssres_stor > sat, None)
        if edited and soilzone_mode == "error":
            raise ValueError("Some parameter values were edited and an error was requested."
                             " See warnings for additional details.")
    if channel_mode in ("warn", "error") and "seg_slope" in pd:
        flat = np.asarray(pd["seg_slope"], dtype=np.float64) < 1e-7
        if flat.any():
            warn(f"seg_slope < 1.0e-7, set to 1.0e-4 at indices:{np.where(flat)[0]}", UserWarning,
                 stacklevel=4)
            if channel_mode == "error":
                raise ValueError("seg_slope parameter values were edited and an error was "
                                 "requested. See warnings for additional details.")


def base_name(name: str) -> str:
    """PRMSRunoffNoDprst -> PRMSRunoff: the process slot in the chain."""
    return name.replace("NoDprst", "")


def _kw(loc):
    return {k: v for k, v in loc.items()
            if k not in ("self", "control", "discretization", "parameters", "__class__", "kw")}


PROCESS_CLASSES = {c.__name__: c for c in (
    PRMSSolarGeometry, PRMSAtmosphere, PRMSCanopy, PRMSSnow, PRMSRunoff, PRMSSoilzone,
    PRMSGroundwater, PRMSChannel, PRMSRunoffNoDprst, PRMSSoilzoneNoDprst, PRMSGroundwaterNoDprst,
    PRMSEt)}

# base/model.py:15-28
PROCESS_ORDER_NHM = ("PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow",
                     "PRMSRunoff", "PRMSRunoffNoDprst", "PRMSSoilzone", "PRMSSoilzoneNoDprst",
                     "PRMSEt", "PRMSGroundwater", "PRMSGroundwaterNoDprst", "PRMSChannel")
