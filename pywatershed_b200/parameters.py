"""Parameters: read-only dict-of-numpy parameter set.

The slice of the reference's Parameters / PrmsParameters API the process chain
uses (pywatershed/base/parameters.py:15-227,
parameters/prms_parameters.py:132-274): `.parameters`, `.dims`, `.coords`,
`.subset()`, `.get_param_values()`, `PrmsParameters.load(param_file)` incl.
the derived `hru_in_to_cf`.  A reference Parameters object works too.
"""
from __future__ import annotations

from types import MappingProxyType

import numpy as np

from .prms_files import PrmsFile

FT2_PER_ACRE = 43560.0
INCHES_PER_FOOT = 12.0


class Parameters:
    def __init__(self, dims: dict = None, coords: dict = None, data_vars: dict = None,
                 metadata: dict = None, validate: bool = True):
        self._dims = dict(dims or {})
        self._coords = {k: _ro(v) for k, v in (coords or {}).items()}
        self._data = {k: _ro(v) for k, v in (data_vars or {}).items()}
        self._metadata = dict(metadata or {})

    @property
    def dims(self):
        return MappingProxyType(self._dims)

    @property
    def dimensions(self):
        return self.dims

    @property
    def coords(self):
        return MappingProxyType(self._coords)

    @property
    def parameters(self):
        return MappingProxyType(self._data)

    @property
    def data_vars(self):
        return self.parameters

    @property
    def variables(self):
        return MappingProxyType({**self._coords, **self._data})

    @property
    def metadata(self):
        return MappingProxyType(self._metadata)

    def get_param_values(self, keys=None):
        if keys is None:
            return self.parameters
        if isinstance(keys, str):
            return self._data[keys]
        return {k: self._data[k] for k in keys}

    def get_dim_values(self, keys=None):
        if keys is None:
            return self.dims
        if isinstance(keys, str):
            return self._dims[keys]
        return {k: self._dims[k] for k in keys}

    def subset(self, keys, **kwargs):
        keys = [k for k in keys if k in self._data]
        return type(self)(dims=self._dims, coords=self._coords,
                          data_vars={k: self._data[k] for k in keys},
                          metadata={k: self._metadata.get(k) for k in keys})

    @staticmethod
    def merge(*params):
        dims, coords, data, meta = {}, {}, {}, {}
        for p in params:
            if p is None:
                continue
            dims.update(dict(p.dims))
            coords.update(dict(p.coords))
            data.update(dict(p.parameters))
            meta.update(dict(getattr(p, "metadata", {})))
        return Parameters(dims=dims, coords=coords, data_vars=data, metadata=meta)

    # ---- netCDF round trip (base/data_model.py:354-420).  Reading goes through
    # netcdf_in.open_netcdf: netCDF4 when it is importable, else this package's own
    # HDF5 reader (netCDF-4 files, what pywatershed writes) or scipy (classic files).
    @classmethod
    def from_netcdf(cls, nc_file, *args, **kwargs):
        from .netcdf_in import open_netcdf

        coords, data, meta = {}, {}, {}
        with open_netcdf(nc_file) as ds:
            dims = dict(ds.dims)
            for name, var in ds.variables.items():
                arr = np.asarray(var.read())
                meta[name] = {"dims": tuple(var.dims)}
                (coords if name in ("nhm_id", "nhm_seg", "doy") or name in dims else data)[name] = arr
        return cls(dims=dims, coords=coords, data_vars=data, metadata=meta)

    def to_netcdf(self, filename, use_xr=False) -> None:
        from .netcdf_out import nc4_module

        nc4 = nc4_module()
        meta = self._metadata
        if nc4 is not None:
            ds = nc4.Dataset(filename, "w", format="NETCDF4")
        else:
            from scipy.io import netcdf_file

            ds = netcdf_file(str(filename), "w", version=2)
        try:
            for k, n in self._dims.items():
                ds.createDimension(k, int(n))
            for name, arr in {**self._coords, **self._data}.items():
                arr = np.asarray(arr)
                dnames = tuple((meta.get(name) or {}).get("dims") or ())
                if len(dnames) != arr.ndim:  # fall back on matching lengths
                    dnames = tuple(next(k for k, n in self._dims.items() if n == m) for m in arr.shape)
                if arr.dtype == np.int64:
                    arr = arr.astype(np.int32)
                v = ds.createVariable(name, arr.dtype.str[1:], dnames)
                v[:] = arr
        finally:
            ds.close()

    def to_dict(self) -> dict:
        """Plain dict name -> ndarray incl. coordinates (what Engine takes)."""
        out = {k: np.asarray(v) for k, v in self._data.items()}
        out.update({k: np.asarray(v) for k, v in self._coords.items()})
        return out


class PrmsParameters(Parameters):
    @staticmethod
    def load(parameter_file) -> "PrmsParameters":
        pf = PrmsFile(parameter_file, "parameter")
        data = dict(pf.data)
        dims = dict(pf.dims)
        dims.setdefault("ndoy", 366)
        dims["nmonth"] = dims.get("nmonths", 12)
        coords = {}
        for cc, dname in (("nhm_id", "nhru"), ("nhm_seg", "nsegment")):
            if cc in data:
                coords[cc] = data.pop(cc)
            elif dname in dims:
                coords[cc] = np.arange(1, dims[dname] + 1)
        # derived parameter, parameters/prms_parameters.py:249-252
        data["hru_in_to_cf"] = data["hru_area"] * FT2_PER_ACRE / (INCHES_PER_FOOT)
        meta = {k: {"dims": v} for k, v in pf.param_dims.items()}
        meta["hru_in_to_cf"] = {"dims": ("nhru",)}
        return PrmsParameters(dims=dims, coords=coords, data_vars=data, metadata=meta)


def _ro(v):
    a = np.array(v, copy=True)
    a.flags.writeable = False
    return a


def as_param_dict(parameters) -> dict:
    """name -> ndarray from our Parameters, a reference Parameters, or a dict."""
    if isinstance(parameters, dict):
        return parameters
    if hasattr(parameters, "to_dict") and isinstance(parameters, Parameters):
        return parameters.to_dict()
    out = {k: np.asarray(v) for k, v in dict(parameters.parameters).items()}
    for k, v in dict(getattr(parameters, "coords", {})).items():
        out.setdefault(k, np.asarray(v))
    return out
