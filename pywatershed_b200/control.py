"""Control: the global clock and options dict of a run.

Mirrors the part of the reference's Control the process chain touches
(pywatershed/base/control.py:80-639): same constructor, `load_prms`, the
`options` whitelist, `advance()` and the current_* calendar properties.
A reference `pywatershed.Control` instance can be passed to Model instead;
only the attributes used here are required (duck typing).
"""
from __future__ import annotations

import pathlib as pl
from warnings import warn

import numpy as np

from .prms_files import PrmsFile

# base/control.py:30-45 plus this package's own knobs
PWS_OPTIONS = (
    "budget_type", "calc_method", "dprst_flag", "input_dir", "load_n_time_batches",
    "netcdf_output_dir", "netcdf_output_var_names", "netcdf_output_separate_files",
    "netcdf_budget_args", "parameter_file", "start_time", "end_time", "time_step",
    "time_step_units", "verbosity", "streamflow_module",
    # pywatershed_b200 additions
    "block_days", "device",
    # "netcdf4": netCDF-4 / HDF5 files like the reference's (zlib 4, shuffle, one chunk per day) from
    # our own writer (hdf5_out.py) when the netCDF4 module is absent; default "classic" (NetCDF-3)
    "netcdf_output_format",
)

# base/control.py:47-72: PRMS legacy option -> pywatershed option
PRMS_TO_PWS = {
    "start_time": "start_time", "end_time": "end_time", "dprst_flag": "dprst_flag",
    "initial_deltat": "time_step", "nhruOutBaseFileName": "netcdf_output_dir",
    "nhruOutVar_names": "netcdf_output_var_names",
    "nsegmentOutBaseFileName": "netcdf_output_dir",
    "nsegmentOutVar_names": "netcdf_output_var_names", "param_file": "parameter_file",
    "print_debug": "verbosity", "strmflow_module": "streamflow_module",
}


class OptsDict(dict):
    """Options dict that refuses unknown keys (base/control.py:623-639)."""

    def __setitem__(self, key, value):
        if key not in PWS_OPTIONS:
            raise NameError(f"'{key}' is not an available control option")
        super().__setitem__(key, value)

    def update(self, other=(), **kw):
        for k, v in dict(other, **kw).items():
            self[k] = v


class Control:
    def __init__(self, start_time: np.datetime64, end_time: np.datetime64,
                 time_step: np.timedelta64, init_time: np.datetime64 = None,
                 options: dict = None, only_warn_invalid: bool = False):
        self.name = "Control"
        if end_time < start_time:
            raise ValueError("end_time < start_time")
        n_times_m1 = (end_time - start_time) / time_step
        if n_times_m1 != int(n_times_m1):
            raise ValueError("time_step does not divide end_time - start_time")
        self._start_time = start_time
        self._end_time = end_time
        self._time_step = time_step
        self._n_times = int(n_times_m1) + 1
        self._init_time = init_time if init_time else self._start_time - time_step
        self._current_time = self._init_time
        self._previous_time = None
        self._itime_step = -1
        opts = OptsDict()
        if options:
            for k, v in options.items():
                try:
                    opts[k] = v
                except NameError:
                    if only_warn_invalid:
                        warn(f"'{k}' is not an available control option", RuntimeWarning)
                    else:
                        raise
        self.options = opts

    # ---- constructors
    @classmethod
    def load_prms(cls, control_file, warn_unused_options: bool = True) -> "Control":
        """PRMS-native ASCII control file (base/control.py:227-311)."""
        raw = PrmsFile(control_file, "control").data
        opts = OptsDict()
        unused = []
        for k, v in raw.items():
            if k in PRMS_TO_PWS:
                pk = PRMS_TO_PWS[k]
                if pk in ("start_time", "end_time", "time_step"):
                    continue
                if pk == "netcdf_output_var_names":
                    opts[pk] = list(opts.get(pk, [])) + [str(x) for x in np.atleast_1d(v)]
                elif pk == "dprst_flag":
                    opts[pk] = bool(np.atleast_1d(v)[0])
                elif pk == "netcdf_output_dir":
                    opts[pk] = pl.Path(str(np.atleast_1d(v)[0])).parent
                elif pk == "verbosity":
                    opts[pk] = int(np.atleast_1d(v)[0])
                else:
                    opts[pk] = np.atleast_1d(v)[0] if np.size(v) == 1 else v
            else:
                unused.append(k)
        if warn_unused_options and unused:
            warn(f"Option(s) unused by pywatershed_b200: {unused}", RuntimeWarning)
        start = _prms_time(raw["start_time"])
        end = _prms_time(raw["end_time"])
        dt_h = float(np.atleast_1d(raw.get("initial_deltat", [24.0]))[0])
        return cls(start, end, np.timedelta64(int(dt_h), "h"), options=opts)

    @classmethod
    def from_yaml(cls, yaml_file) -> "Control":
        """Control from a yaml file (base/control.py:549-610): start_time, end_time,
        time_step, time_step_units at the top level, every other key is an option;
        input_dir is taken relative to the yaml file and must exist."""
        import yaml

        yaml_file = pl.Path(yaml_file)
        with yaml_file.open("r") as fh:
            d = yaml.load(fh, Loader=yaml.Loader)
        start = np.datetime64(d.pop("start_time"))
        end = np.datetime64(d.pop("end_time"))
        step = np.timedelta64(int(d.pop("time_step")), d.pop("time_step_units"))
        if "input_dir" in d:
            path = pl.Path(d["input_dir"])
            d["input_dir"] = path if path.is_absolute() else (yaml_file.parent / path).resolve()
            if not d["input_dir"].exists():
                raise FileNotFoundError(f"input_dir {d['input_dir']} does not exist")
        return cls(start, end, step, options=d)

    def to_dict(self, deep_copy: bool = True) -> dict:
        """base/control.py:490-516."""
        from copy import deepcopy

        step_h = int(self._time_step / np.timedelta64(1, "h"))
        opts = deepcopy(dict(self.options)) if deep_copy else dict(self.options)
        return {"start_time": str(self._start_time), "end_time": str(self._end_time),
                "time_step": step_h, "time_step_units": "h", "options": opts}

    def to_yaml(self, yaml_file) -> None:
        """Options flattened to the top level next to the four time keys
        (base/control.py:518-547); `from_yaml` reads it back."""
        import yaml

        d = self.to_dict()
        opts = d.pop("options")
        for k, v in opts.items():
            if k in d:
                raise ValueError("Control option keys collide with non-option keys")
            d[k] = str(v) if isinstance(v, pl.Path) else v
        with open(yaml_file, "w") as fh:
            yaml.dump(d, fh)

    # ---- clock
    @property
    def current_time(self):
        return self._current_time

    @property
    def current_datetime(self):
        return self._current_time.astype("datetime64[s]").astype(object)

    @property
    def previous_time(self):
        return self._previous_time

    @property
    def itime_step(self):
        return self._itime_step

    @property
    def time_step(self):
        return self._time_step

    @property
    def init_time(self):
        return self._init_time

    @property
    def start_time(self):
        return self._start_time

    @property
    def end_time(self):
        return self._end_time

    @property
    def n_times(self):
        return self._n_times

    @property
    def time_step_seconds(self):
        return self._time_step / np.timedelta64(1, "s")

    @property
    def time_step_days(self):
        return self._time_step / np.timedelta64(1, "D")

    def _cal(self, t):
        from .engine import calendar

        return calendar(t, 1)[0]

    @property
    def current_doy(self):
        return int(self._cal(self._current_time)[0])

    @property
    def current_dowy(self):
        return int(self._cal(self._current_time)[1])

    @property
    def current_month(self):
        return int(self._cal(self._current_time)[2])

    @property
    def current_year(self):
        return int(self._current_time.astype("datetime64[Y]").astype(int) + 1970)

    @property
    def start_doy(self):
        return int(self._cal(self._start_time)[0])

    @property
    def start_month(self):
        return int(self._cal(self._start_time)[2])

    def advance(self, n: int = 1):
        """base/control.py:439-453; `n` steps at once (a time-blocked Model.run)."""
        for _ in range(int(n)):
            if self._current_time == self._end_time:
                raise ValueError("End of time reached")
            self._itime_step += 1
            self._previous_time = self._current_time
            self._current_time = self._current_time + self._time_step

    def edit_end_time(self, new_end_time: np.datetime64):
        self._end_time = new_end_time
        self._n_times = int((self._end_time - self._start_time) / self._time_step) + 1

    def edit_n_time_steps(self, new_n_time_steps: int):
        self._n_times = new_n_time_steps
        self._end_time = self._start_time + (self._n_times - 1) * self._time_step

    def __getitem__(self, key):
        return getattr(self, key)


def _prms_time(v) -> np.datetime64:
    y, m, d, hh, mm, ss = [int(x) for x in np.atleast_1d(v)[:6]]
    return np.datetime64(f"{y:04d}-{m:02d}-{d:02d}T{hh:02d}:{mm:02d}:{ss:02d}")
