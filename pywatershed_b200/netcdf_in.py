"""Reading the reference's netCDF inputs (utils/netcdf_utils.py:31-353,
base/data_model.py:354-420): parameter files and [time, nhm_id] forcings.

Backends, in this order: the netCDF4 module when it is installed; this
package's own minimal HDF5 reader (hdf5_min.py) for netCDF-4 files; scipy's
reader for classic netCDF-3 files.  All three present the same view:

    with open_netcdf(path) as ds:
        ds.dims                      # {name: length}
        ds.variables[name].dims      # tuple of dimension names
        ds.variables[name].attrs     # dict
        ds.variables[name].read()    # ndarray
"""
from __future__ import annotations

import numpy as np


class _Var:
    def __init__(self, dims, attrs, reader, dtype, shape):
        self.dims = tuple(dims)
        self.attrs = attrs
        self.read = reader
        self.dtype = dtype
        self.shape = tuple(shape)


class _Dataset:
    def __init__(self, dims, variables, closer):
        self.dims = dims
        self.variables = variables
        self._closer = closer

    def close(self):
        self._closer()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _decode(v):
    return v.decode() if isinstance(v, bytes) else v


def open_netcdf(path) -> _Dataset:
    from .netcdf_out import nc4_module

    nc4 = nc4_module()
    if nc4 is not None:
        ds = nc4.Dataset(path)
        dims = {k: len(v) for k, v in ds.dimensions.items()}
        variables = {}
        for name, var in ds.variables.items():
            var.set_auto_mask(False)
            variables[name] = _Var(var.dimensions, {a: var.getncattr(a) for a in var.ncattrs()},
                                   (lambda v=var: np.asarray(v[:])), var.dtype, var.shape)
        return _Dataset(dims, variables, ds.close)
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic[:3] == b"CDF":
        from scipy.io import netcdf_file

        ds = netcdf_file(str(path), "r", mmap=False)
        dims = {k: (v if v is not None else 0) for k, v in ds.dimensions.items()}
        variables = {}
        for name, var in ds.variables.items():
            attrs = {k: _decode(v) for k, v in var._attributes.items()}
            variables[name] = _Var(var.dimensions, attrs, (lambda v=var: np.array(v[:])),
                                   var.data.dtype, var.shape)
            if var.isrec and var.dimensions:
                dims[var.dimensions[0]] = var.shape[0]
        return _Dataset(dims, variables, ds.close)
    from . import hdf5_min

    f = hdf5_min.File(path)
    dims, variables = {}, {}
    dsets = f.datasets()
    for name, d in dsets.items():
        if d.attrs.get("CLASS") == "DIMENSION_SCALE" and len(d.shape or ()) == 1:
            dims[name] = int(d.shape[0])
    for name, d in dsets.items():
        if d.is_bare_dimension:
            continue
        if d.dt.kind not in ("num", "str"):
            continue  # variable-length strings etc.: nothing on this path needs them
        dn = d.dims
        for dname, n in zip(dn, d.shape or ()):
            # an unlimited dimension's scale may be shorter than the variables on it
            dims[dname] = max(dims.get(dname, 0), int(n))
        attrs = {k: v for k, v in d.attrs.items()
                 if k not in ("DIMENSION_LIST", "REFERENCE_LIST", "CLASS", "NAME", "_Netcdf4Dimid",
                              "_Netcdf4Coordinates", "_nc3_strict")}
        variables[name] = _Var(dn, attrs, d.read, d.dtype, d.shape or ())
    return _Dataset(dims, variables, f.close)


def decode_times(values, units: str) -> np.ndarray:
    """CF time -> datetime64[s] for the '<unit> since <reference>' encodings
    pywatershed files use (standard calendar)."""
    unit, _, ref = _decode(units).partition(" since ")
    ref = ref.strip()
    if " " in ref:
        date, _, clock = ref.partition(" ")
        ref = date + "T" + clock.split(" ")[0]
    ref = np.datetime64(ref.rstrip("Z"), "s")
    sec = {"days": 86400.0, "day": 86400.0, "hours": 3600.0, "hour": 3600.0, "minutes": 60.0,
           "seconds": 1.0, "second": 1.0}[unit.strip().lower()]
    off = np.round(np.asarray(values, dtype=np.float64) * sec).astype("int64")
    return ref + off.astype("timedelta64[s]")


def read_time_variable(path, name, keep_f32=False):
    """[time, unit] variable `name` as float64 (or, with `keep_f32`, as the
    float32 it is stored as) + its times (datetime64[s]).  The time coordinate is
    `time` (or `datetime`, as in the reference's test_data/hru_1/cbh.nc)."""
    with open_netcdf(path) as ds:
        if name not in ds.variables:
            raise KeyError(f"{path}: no variable '{name}' (has {sorted(ds.variables)})")
        var = ds.variables[name]
        tname = next((t for t in ("time", "datetime") if t in ds.variables
                      and ds.variables[t].shape and ds.variables[t].shape[0] == var.shape[0]
                      and "units" in ds.variables[t].attrs), None)
        if tname is None:
            raise ValueError(f"{path}: no time coordinate with units for '{name}'")
        t = ds.variables[tname]
        times = decode_times(t.read(), t.attrs["units"])
        data = np.asarray(var.read())
        if not (keep_f32 and data.dtype == np.float32):
            data = data.astype(np.float64)
        return times, data
