"""netCDF output with the reference's layout (pywatershed/utils/netcdf_utils.py:356-697,
base/process.py:463-592, base/budget.py:661-748):

  <output_dir>/<var>.nc          separate_files=True (default)
  <output_dir>/<Process>.nc      separate_files=False
  <output_dir>/<Process>_budget.nc   inputs_sum / outputs_sum / storage_changes_sum (+ every term
                                     with budget_args={'write_individual_vars': True})

dims `time` (unlimited) x `nhm_id` | `nhm_seg`; coordinate variables
nhm_id / nhm_seg (i4) and time (f4, "days since 1970-01-01 00:00:00");
variable attributes from the metadata; global attributes
Description="pywatershed output data" and "process class".
SolarGeometry / Atmosphere variables are all-time arrays written once per block
with the same dims (their time axis is `doy` for the soltab tables).

Backend: netCDF4 (zlib level 4, chunks (1, n), exactly the reference's files)
when it is importable; otherwise our own streaming writer of the classic
format (cdf_out.py: NetCDF-3 64-bit offset with the same names, dims, dtypes
and attributes but no compression; a block of days is one contiguous write and
nothing is held in memory).  Whole blocks are written with
`var[t0:t1, :] = block`; the files of a block are written by a few threads.
With the control option `netcdf_output_format: netcdf4` (and no netCDF4 module) the
files are netCDF-4 / HDF5 from hdf5_out.H5Writer instead: zlib 4, shuffle, one chunk
per day, dimension scales -- the reference's form, at zlib's pace.
"""
from __future__ import annotations

import pathlib as pl

import numpy as np

from ._process_meta import VARIABLE_META



def nc4_module():
    """The real netCDF4 module, or None (absent, or a test stub standing in
    for it in sys.modules)."""
    try:
        import netCDF4 as m
    except Exception:  # noqa: BLE001
        return None
    return m if isinstance(getattr(m, "Dataset", None), type) else None


_nc4 = nc4_module()
# without the netCDF4 module: "classic" = cdf_out.Cdf2Writer (NetCDF-3 64-bit offset,
# uncompressed, the fast default), "netcdf4" = hdf5_out.H5Writer (netCDF-4 / HDF5 as the
# reference writes it: zlib 4, shuffle, one chunk per day); control option netcdf_output_format
_FORMAT = "classic"

_NC_TYPES = {"float64": "f8", "int32": "i4", "bool": "i4", "int64": "i8"}
# netCDF4.default_fillvals, what the reference passes as fill_value (utils/netcdf_utils.py:607)
_FILL = {"f8": 9.969209968386869e36, "f4": 9.969209968386869e36, "i4": -2147483647, "i8": -9223372036854775806}


def _var_attrs(name):
    """Every scalar metadata item of a variable (the reference copies all of
    them, utils/netcdf_utils.py:613-616: desc, units, type, var_category, ...;
    `dims` is a list and is written as the reference's netCDF4 would, an array of
    strings is not representable in the classic format, so it is joined)."""
    m = VARIABLE_META.get(name, {})
    out = {}
    for k, v in m.items():
        if isinstance(v, dict):
            continue
        out[k] = " ".join(v) if isinstance(v, (list, tuple)) else v
    return out
_THREAD_BYTES = 64 << 20  # blocks at least this large are written by one thread per file


class _File:
    """One netCDF file holding one or more [time, unit] variables."""

    def __init__(self, path: pl.Path, var_names, coord_name, coord_vals, process_name,
                 time_name="time", n_doy=None, global_attrs=None):
        self.path = path
        self.vars = list(var_names)
        self.coord_name = coord_name
        self.time_name = time_name
        path.parent.mkdir(parents=True, exist_ok=True)
        if _nc4 is not None:
            ds = _nc4.Dataset(path, "w", format="NETCDF4")
            ds.setncattr("Description", "pywatershed output data")
            ds.setncattr("process class", process_name)
            for k, val in (global_attrs or {}).items():
                ds.setncattr(k, val)
            ds.createDimension(time_name, n_doy)
            ds.createDimension(coord_name, len(coord_vals))
            # utils/netcdf_utils.py:474-489: time is f4 "days since ...", doy is i4 "Day of year"
            tv = ds.createVariable(time_name, "f4" if time_name == "time" else "i4", (time_name,))
            tv.units = "days since 1970-01-01 00:00:00" if time_name == "time" else "Day of year"
            cv = ds.createVariable(coord_name, "i4", (coord_name,))
            cv[:] = np.asarray(coord_vals, dtype=np.int32)
            self.v = {}
            for name in self.vars:
                m = VARIABLE_META.get(name, {})
                kind = _NC_TYPES[m.get("type", "float64")]
                v = ds.createVariable(name, kind, (time_name, coord_name), fill_value=_FILL[kind],
                                      zlib=True, complevel=4, chunksizes=(1, len(coord_vals)))
                for k, val in _var_attrs(name).items():
                    v.setncattr(k, val)
                self.v[name] = v
        else:
            if _FORMAT == "netcdf4":
                from .hdf5_out import H5Writer as Cdf2Writer
            else:
                from .cdf_out import Cdf2Writer

            ds = Cdf2Writer(path)
            ds.setncattr("Description", "pywatershed output data")
            ds.setncattr("process class", process_name)
            for k, val in (global_attrs or {}).items():
                ds.setncattr(k, val)
            ds.createDimension(time_name, n_doy)
            ds.createDimension(coord_name, len(coord_vals))
            tatts = {"units": "days since 1970-01-01 00:00:00" if time_name == "time" else "Day of year"}
            tv = ds.createVariable(time_name, "f4" if time_name == "time" else "i4", (time_name,), **tatts)
            self._cv = ds.createVariable(coord_name, "i4", (coord_name,))
            self._coord_vals = np.asarray(coord_vals, dtype=np.int32)
            self.v = {}
            for name in self.vars:
                m = VARIABLE_META.get(name, {})
                kind = _NC_TYPES[m.get("type", "float64")]
                atts = _var_attrs(name)
                atts["_FillValue"] = np.array(_FILL[kind], dtype={"f8": np.float64, "i4": np.int32}[kind])
                self.v[name] = ds.createVariable(name, kind, (time_name, coord_name))
                for k, val in atts.items():  # ("dims" is a metadata key: not a keyword argument)
                    self.v[name].setncattr(k, val)
        self.ds = ds
        self.tv = tv

    def group_name(self, group, name):
        """Name of variable `name` of netCDF-4 group `group` (created on demand);
        the classic format has no groups: `<group>__<name>` there."""
        if _nc4 is not None:
            if group not in self.ds.groups:
                self.ds.createGroup(group)
            return f"{group}/{name}"
        return f"{group}__{name}"

    def create_variable(self, name, kind, meta=None):
        """One more [time, unit] variable (before the first write)."""
        atts = {}
        for k, val in (meta or {}).items():
            if val is None or isinstance(val, dict):
                continue
            atts[k] = " ".join(val) if isinstance(val, (list, tuple)) else val
        if _nc4 is not None:
            v = self.ds.createVariable(name, kind, (self.time_name, self.coord_name))
        else:
            v = self.v[name] = self.ds.createVariable(name, kind, (self.time_name, self.coord_name))
        for k, val in atts.items():
            v.setncattr(k, val)
        return v

    def write(self, t0, t1, times_days, blocks: dict):
        """Rows [t0, t1) of the time axis and of the variables in `blocks`."""
        tdt = np.float32 if self.time_name == "time" else np.int32
        if _nc4 is not None:
            self.tv[t0:t1] = np.asarray(times_days, dtype=tdt)
            for name, arr in blocks.items():
                if name in self.v:
                    a = np.asarray(arr)
                    if a.dtype == np.bool_:
                        a = a.astype(np.int32)
                    self.v[name][t0:t1, :] = a
            return
        if self._coord_vals is not None:  # the header is complete now: variables are all declared
            self._cv[:] = self._coord_vals
            self._coord_vals = None
        arrays = {self.time_name: np.asarray(times_days, dtype=tdt)}
        arrays.update({k: np.asarray(a) for k, a in blocks.items() if k in self.v})
        if self.ds.rec_dim is not None:
            self.ds.write_record_block(t0, t1, arrays)  # one contiguous write when all variables are there
        else:
            for k, a in arrays.items():
                self.ds.vars[k][t0:t1] = a

    def close(self):
        if _nc4 is None and self._coord_vals is not None:
            self._cv[:] = self._coord_vals
            self._coord_vals = None
        self.ds.close()


class ModelNetcdfOutput:
    def __init__(self, model, output_dir: pl.Path, separate_files: bool, output_vars, budget_args):
        global _FORMAT
        fmt = (getattr(model.control, "options", {}) or {}).get("netcdf_output_format")
        if fmt is not None:
            if fmt not in ("classic", "netcdf4"):
                raise ValueError("netcdf_output_format must be 'classic' or 'netcdf4'")
            _FORMAT = fmt
        self.model = model
        self.dir = pl.Path(output_dir)
        self.files = []
        self.budget_files = {}
        pd = model._param_dict
        nhm_id = np.asarray(pd.get("nhm_id", np.arange(1, model._engine.nhru + 1)))
        nhm_seg = np.asarray(pd.get("nhm_seg", np.arange(1, max(model._engine.nseg, 1) + 1)))
        self._written_static = False
        for pname, proc in model.processes.items():
            names = [v for v in proc.get_variables() if output_vars is None or v in output_vars]
            if not names:
                continue
            if any(v not in VARIABLE_META for v in names):
                # a FlowGraph / exchange behind the column (flow_graph.py): its variables are
                # per node (coordinate node_coord, base/flow_graph.py:542-567) or per HRU
                for v in names:
                    n = int(np.asarray(proc[v]).shape[0])
                    per_hru = n == len(nhm_id) and not hasattr(proc, "nnodes") or (
                        hasattr(proc, "nhru") and n == getattr(proc, "nhru") and n != getattr(proc, "nnodes", -1))
                    self.files.append((pname, False, _File(
                        self.dir / (f"{v}.nc" if separate_files else f"{pname}_{v}.nc"), [v],
                        "nhm_id" if per_hru else "node_coord", nhm_id if per_hru else np.arange(n), pname)))
                continue
            doy = pname == "PRMSSolarGeometry"
            groups = [[v] for v in names] if separate_files else [names]
            for g in groups:
                seg = "nsegment" in VARIABLE_META[g[0]]["dims"]
                same = [v for v in g if ("nsegment" in VARIABLE_META[v]["dims"]) == seg]
                rest = [v for v in g if v not in same]
                for part in (same, rest):
                    if not part:
                        continue
                    pseg = "nsegment" in VARIABLE_META[part[0]]["dims"]
                    fname = f"{part[0]}.nc" if separate_files else (
                        f"{pname}.nc" if part is same else f"{pname}_seg.nc")
                    self.files.append((pname, doy, _File(
                        self.dir / fname, part, "nhm_seg" if pseg else "nhm_id",
                        nhm_seg if pseg else nhm_id, pname,
                        time_name="doy" if doy else "time", n_doy=366 if doy else None)))
            b = getattr(proc, "budget", None)
            if b is not None:
                bf = _BudgetFile(self.dir / f"{pname}_budget.nc", pname, b, nhm_id, **(budget_args or {}))
                if bf.v:
                    self.budget_files[pname] = bf

    def write_days(self, model, t0, t1, blk=None, b0=None):
        """Write days [t0, t1) of a block (default: the model's current one)."""
        if blk is None:
            blk, b0 = model._block, model._block_t0
        start = model.control.start_time.astype("datetime64[D]")
        epoch = np.datetime64("1970-01-01", "D")
        tdays = (start - epoch).astype(int) + np.arange(t0, t1)
        jobs = []
        for pname, doy, f in self.files:
            if doy:
                if not self._written_static:
                    sg = model.processes[pname]
                    jobs.append((f.write, (0, 366, np.arange(1, 367), {v: sg[v].data for v in f.vars})))
                continue
            jobs.append((f.write, (t0, t1, tdays,
                                   {v: blk[v][t0 - b0:t1 - b0] for v in f.vars if v in blk})))
        for pname, bf in self.budget_files.items():
            jobs.append((bf.write, (t0, t1, tdays, blk, t0 - b0, t1 - b0)))
        self._written_static = True
        # Each job touches one file.  The conversion to the file's byte order and
        # the write itself release the GIL, so for blocks worth the trouble a few
        # threads write the files side by side.
        nbytes = sum(getattr(x, "nbytes", 0) for _, a in jobs if isinstance(a[3], dict)
                     for x in a[3].values())
        if len(jobs) > 1 and nbytes >= _THREAD_BYTES and _nc4 is None:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(min(8, len(jobs))) as pool:
                for fut in [pool.submit(fn, *a) for fn, a in jobs]:
                    fut.result()
        else:
            for fn, a in jobs:
                fn(*a)

    def close(self):
        for _, _, f in self.files:
            f.close()
        for bf in self.budget_files.values():
            bf.close()


class _BudgetFile:
    """<Process>_budget.nc (base/budget.py:661-748): the component sums per unit
    (or per domain for a global-basis budget) named by `write_sum_vars` (True:
    all three, a list: those, False / None: none) and, with
    `write_individual_vars`, every term of every component.  The reference puts
    a component's terms into a netCDF-4 group of its name (`inputs/net_rain`);
    NetCDF-3 has no groups, so without the netCDF4 module they are variables
    named `<component>__<term>`."""

    def __init__(self, path, pname, budget, nhm_id, write_sum_vars=True, write_individual_vars=False,
                 **_ignored):
        self.budget = budget
        self.glob = budget.basis == "global"
        if write_sum_vars is True:
            sums = list(budget.output_vars_desc)
        elif isinstance(write_sum_vars, (list, tuple)):
            sums = [v for v in budget.output_vars_desc if v in write_sum_vars]
        elif write_sum_vars in (False, None):
            sums = []
        else:
            raise ValueError(f"Unexpected value of write_sum_vars: {write_sum_vars}")
        self.v, self.term_vars = {}, {}
        self.f = None
        if not sums and not write_individual_vars:
            import warnings

            warnings.warn(f"Budget for {pname} has no requested output, no {pname}_budget.nc is written")
            return
        coord = np.array([0]) if self.glob else nhm_id
        attrs = {"Description": f"pywatershed ({budget.basis}) budget for {budget.description}",
                 "Budget basis": f"{budget.basis} (unit or global)"}
        for c in budget.components:
            attrs[c] = "[" + ", ".join(budget.terms[c]) + "]"
        self.f = _File(path, [], "one" if self.glob else "nhm_id", coord, pname, global_attrs=attrs)
        for name in sums:
            self.v[name] = self.f.create_variable(name, "f8", meta=budget.meta.get(name))
        if write_individual_vars:
            for c in budget.components:
                for t in budget.terms[c]:
                    key = self.f.group_name(c, t)
                    self.term_vars[key] = (c, t)
                    self.v[key] = self.f.create_variable(key, "f8", meta=budget.meta.get(t))
        self.f.v = dict(self.v)

    def write(self, t0, t1, tdays, blk, j0, j1):
        out = {}
        for name in self.v:
            if name in self.term_vars:
                t = self.term_vars[name][1]
                if t not in blk:
                    continue
                a = np.asarray(blk[t][j0:j1], dtype=np.float64)
                out[name] = a.sum(axis=1)[:, None] if self.glob else a
                continue
            terms = [np.asarray(blk[v][j0:j1], dtype=np.float64) for v in self.budget.terms[name[:-4]]
                     if v in blk]
            if not terms:
                continue
            out[name] = sum(t.sum(axis=1) for t in terms)[:, None] if self.glob else sum(terms)
        self.f.write(t0, t1, tdays, out)

    def close(self):
        if self.f is not None:
            self.f.close()
