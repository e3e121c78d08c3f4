"""Import and drive the UNMODIFIED pywatershed reference (numpy calc_method).

TEST INFRASTRUCTURE ONLY.  This module runs only in the build container where
`/root/reference` is mounted; it is used by `oracle/gen_golden.py` to produce
the committed fixtures under `tests/golden/` and by a few optional `not gpu`
tests that are skipped when the reference is absent.  Nothing in the product
package (`pywatershed_b200/`) imports it.

The reference imports netCDF4 / xarray / cftime / ... at module import time;
those are absent here, and the ASCII (PRMS-native) input path does not need
them, so permissive stub modules are inserted first.
"""
from __future__ import annotations

import pathlib as pl
import sys
import types

import numpy as np

REF_ROOT = pl.Path("/root/reference")
DRB_DIR = REF_ROOT / "pywatershed" / "data" / "drb_2yr"

_STUB_NAMES = (
    "netCDF4", "xarray", "cftime", "epiweeks", "flopy", "geopandas",
    "contextily", "xmltodict", "xyzservices", "matplotlib", "pint",
    "shapely", "h5py", "pydot", "folium", "hvplot", "cartopy",
)


class _Stub(types.ModuleType):
    """A module that answers every attribute with another stub."""

    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(f"{self.__name__}.{name}")
        sys.modules[child.__name__] = child
        setattr(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return self

    def __mro_entries__(self, bases):
        return (object,)


class _StubFinder:
    """Meta-path finder: any submodule of a stubbed package is a stub too."""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        import importlib.machinery

        if fullname.split(".")[0] in _STUB_NAMES and isinstance(
            sys.modules.get(fullname.split(".")[0]), _Stub
        ):
            return importlib.machinery.ModuleSpec(fullname, _StubFinder, is_package=True)
        return None

    @staticmethod
    def create_module(spec):
        return _Stub(spec.name)

    @staticmethod
    def exec_module(module):
        return None


def reference_available() -> bool:
    return (REF_ROOT / "pywatershed" / "__init__.py").exists()


def import_reference():
    """Return the reference `pywatershed` module (imported under stubs)."""
    if "pywatershed" in sys.modules and not isinstance(
        sys.modules["pywatershed"], _Stub
    ):
        return sys.modules["pywatershed"]
    for name in _STUB_NAMES:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    if not any(ff is _StubFinder for ff in sys.meta_path):
        sys.meta_path.append(_StubFinder)
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    import pywatershed  # noqa: E402

    return pywatershed


def make_mem_adapter(pws, name, data, times, control):
    """An Adapter serving an in-memory [ntime, nhru] forcing array."""

    class _NcRead:
        pass

    class MemAdapter(pws.base.adapter.Adapter):
        def __init__(self):
            super().__init__(name)
            self.name = "MemAdapter"
            self._fname = f"<memory:{name}>"
            self.control = control
            self._data = np.ascontiguousarray(data, dtype=np.float64)
            self._nc_read = _NcRead()
            self._nc_read.times = times
            self.time = times
            self._current_value = np.full(data.shape[1], np.nan)

        @property
        def data(self):
            return self._data

        def advance(self):
            self._current_value[:] = self._data[self.control.itime_step, :]

    return MemAdapter()


NHM_PROCESS_NAMES = (
    "PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow",
    "PRMSRunoff", "PRMSSoilzone", "PRMSGroundwater", "PRMSChannel",
)


def load_drb(pws):
    """Control, parameters and forcings of the reference's drb_2yr domain."""
    control = pws.Control.load_prms(
        DRB_DIR / "nhm.control", warn_unused_options=False
    )
    params = pws.parameters.PrmsParameters.load(DRB_DIR / "myparam.param")
    from pywatershed.utils.cbh_utils import cbh_files_to_np_dict

    forcings = {}
    times = None
    for var in ("prcp", "tmax", "tmin"):
        dd = cbh_files_to_np_dict({var: DRB_DIR / f"{var}.cbh"}, params)
        forcings[var] = np.asarray(dd[var], dtype=np.float64)
        times = dd["datetime"] if "datetime" in dd else dd.get("time")
    return control, params, forcings, times


def build_model(pws, control, params, forcings, times, process_names=NHM_PROCESS_NAMES,
                calc_method="numpy", budget_type="error"):
    for key in ("netcdf_output_dir", "netcdf_output_var_names"):
        if key in control.options:
            del control.options[key]
    control.options["input_dir"] = "/tmp"
    control.options["budget_type"] = budget_type
    control.options["calc_method"] = calc_method
    procs = [getattr(pws, nn) for nn in process_names]
    model = pws.Model(procs, control=control, parameters=params, find_input_files=False)
    for proc in model.processes.values():
        for var in ("prcp", "tmax", "tmin"):
            if var in proc.get_inputs():
                proc.set_input_to_adapter(
                    var, make_mem_adapter(pws, var, forcings[var], times, control)
                )
    model._found_input_files = True
    return model


def run_and_dump(model, n_days=None, hru_subset=None):
    """Run `model` day by day and collect every public variable of every
    process after each day's `calculate()` (the values a user would read
    between steps, cf. autotest/test_prms_below_snow.py:279-303).

    Returns {var: ndarray[ntime, n]} (+ 3 soltab arrays [366, nhru]) and
    {process: ndarray[ntime] of budget._zero_sum flags}.
    """
    control = model.control
    if n_days is None:
        n_days = control.n_times
    out = {}
    shapes = {}
    for pname, proc in model.processes.items():
        for var in proc.get_variables():
            val = proc[var]
            if hasattr(val, "data") and hasattr(val, "current"):
                continue  # TimeseriesArray: collected after the run
            shapes[var] = (pname, val.shape, val.dtype)
    for var, (pname, shp, dt) in shapes.items():
        out[var] = np.empty((n_days,) + shp, dtype=dt)
    budgets = {}
    for istep in range(n_days):
        model.advance()
        model.calculate()
        for var, (pname, shp, dt) in shapes.items():
            out[var][istep] = model.processes[pname][var]
    for pname, proc in model.processes.items():
        for var in proc.get_variables():
            val = proc[var]
            if hasattr(val, "data") and hasattr(val, "current"):
                arr = np.asarray(val.data)
                out[var] = arr[:n_days] if arr.shape[0] != 366 else arr
    return out
