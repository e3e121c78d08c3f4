"""Readers for PRMS-native ASCII inputs: control file, parameter file, CBH.

Formats as parsed by the reference (pywatershed/utils/prms5_file_util.py:60-403,
utils/cbh_utils.py:32-122): `####`-separated records; a control record is
name / count / type-code / values, a parameter record is name / ndims /
dimension names / count / type-code / values (Fortran order for 2-D), type
codes 1 int, 2 float32, 3 float64, 4 string.
"""
from __future__ import annotations

import pathlib as pl

import numpy as np

_TYPES = {1: np.int64, 2: np.float64, 3: np.float64, 4: str}


class PrmsFile:
    def __init__(self, path, file_type: str):
        self.path = pl.Path(path)
        self.dims: dict[str, int] = {}
        self.data: dict[str, np.ndarray] = {}
        self.param_dims: dict[str, tuple] = {}
        lines = self.path.read_text().splitlines()
        if file_type == "control":
            self._parse_control(lines)
        elif file_type == "parameter":
            self._parse_parameter(lines)
        else:
            raise ValueError(file_type)

    @staticmethod
    def _records(lines):
        rec = []
        for ln in lines:
            s = ln.strip()
            if s.startswith("####"):
                if rec:
                    yield rec
                rec = []
            else:
                rec.append(s)
        if rec:
            yield rec

    def _parse_control(self, lines):
        first = next(i for i, ln in enumerate(lines) if ln.strip().startswith("####"))
        for rec in self._records(lines[first:]):
            rec = [r for r in rec if r != ""]
            if len(rec) < 3:
                continue
            name, n, tcode = rec[0].split()[0], int(rec[1]), int(rec[2])
            vals = rec[3:3 + n]
            typ = _TYPES[tcode]
            self.data[name] = np.array(vals, dtype=object).astype(str) if typ is str else \
                np.array([typ(float(v)) if typ is not np.int64 else int(float(v)) for v in vals])

    def _parse_parameter(self, lines):
        i_dim = next(i for i, ln in enumerate(lines) if ln.strip().startswith("** Dimensions **"))
        i_par = next(i for i, ln in enumerate(lines) if ln.strip().startswith("** Parameters **"))
        for rec in self._records(lines[i_dim + 1:i_par]):
            self.dims[rec[0].split()[0]] = int(rec[1])
        for rec in self._records(lines[i_par + 1:]):
            name = rec[0].split()[0]
            ndim = int(rec[1])
            dnames = tuple(rec[2:2 + ndim])
            n = int(rec[2 + ndim])
            tcode = int(rec[3 + ndim])
            vals = rec[4 + ndim:4 + ndim + n]
            typ = _TYPES[tcode]
            arr = np.array(vals, dtype=str) if typ is str else np.array(vals, dtype=typ)
            if ndim == 2:
                # PRMS stores (nhru, nmonths) in Fortran order; the reference
                # presents them as [nmonth, nhru] (prms5_file_util.py:351-403)
                d0, d1 = self.dims[dnames[0]], self.dims[dnames[1]]
                arr = arr.reshape((d1, d0))
                dnames = (dnames[1], dnames[0])
            self.data[name] = arr
            self.param_dims[name] = dnames


def read_cbh(path, nhru: int | None = None):
    """CBH ASCII -> (times datetime64[s] [ntime], values float64 [ntime, nhru])
    (utils/cbh_utils.py:32-122)."""
    path = pl.Path(path)
    with open(path) as f:
        n_in_file = None
        for ln in f:
            if ln.startswith("####"):
                break
            parts = ln.split()
            if len(parts) == 2 and parts[1].isdigit():
                n_in_file = int(parts[1])
        try:  # pandas' C parser reads a CONUS-sized CBH file ~20x faster than loadtxt
            import pandas as pd

            raw = pd.read_csv(f, sep=r"\s+", header=None, dtype=np.float64, engine="c").to_numpy()
        except ImportError:
            raw = np.loadtxt(f, dtype=np.float64, ndmin=2)
    ymdhms = raw[:, :6].astype(int)
    vals = np.ascontiguousarray(raw[:, 6:])
    if n_in_file is not None and vals.shape[1] != n_in_file:
        raise ValueError(f"{path}: header says {n_in_file} columns, found {vals.shape[1]}")
    if nhru is not None and vals.shape[1] != nhru:
        raise ValueError(f"{path}: {vals.shape[1]} columns but nhru = {nhru}")
    times = np.array([np.datetime64(f"{y:04d}-{m:02d}-{d:02d}T{hh:02d}:{mm:02d}:{ss:02d}")
                      for y, m, d, hh, mm, ss in ymdhms])
    return times, vals
